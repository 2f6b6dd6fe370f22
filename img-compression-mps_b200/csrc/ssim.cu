// K9: SSIM exactly as the reference defines it (utils/metrics.py:11-129 on top of
// scikit-image's uniform-window structural_similarity).
//
// The reference's 3-D / 4-D SSIM is NOT a 3-D stencil: it is the mean over the three
// slicing axes of the mean over 2-D slices, each slice with its own data range and a
// 7x7 box window (min(7, smallest side) forced odd), the second argument clipped at 0,
// sample covariance, borders cropped.  A "family" below is one slicing axis; a 4-D
// array contributes frames x slices per family.
//
// Pass 1: per-slice data range.  Pass 2: tiles of TILE_H x TILE_W interior pixels,
// separable box sums of x, y, xx, yy, xy in float64 from a shared-memory patch, one
// partial sum per tile.  Pass 3: fixed-order reduction (deterministic).
//
// float32 inputs with the 7 x 7 window take the STREAMING kernels further down instead of the tile
// kernels (context option `ssim_exact` = 1 forces the tile kernels): a warp walks down the rows of a
// strip, every thread keeps the last seven rows of its column in registers, window sums are float64
// (products and sums of float32 values are exact there, so the cancellation in E[xx] - E[x]^2 costs
// nothing), only the final SSIM formula is evaluated in float32 on range-normalised moments.
// No shared memory, no index tables; agreement with the float64 formula ~1e-7.
#include "common.cuh"

namespace ndmps {

struct SliceFamily {
    int64_t S;        // slices (times frames)
    int64_t H, W;     // slice extent
    int64_t sh, sw;   // element strides of the slice's rows / columns
    int64_t nT;       // frames (1 for 2-D / 3-D)
    int64_t s_axis;   // stride of the slicing axis
    int64_t s_t;      // stride of the frame axis
    int win;          // window (odd)
};

constexpr int TILE_W = 32, TILE_H = 32, MAXWIN = 7;
constexpr int PATCH_W = TILE_W + MAXWIN - 1, PATCH_H = TILE_H + MAXWIN - 1;

__device__ __forceinline__ int64_t slice_base(const SliceFamily& f, int64_t s) {
    int64_t i = s / f.nT, t = s - i * f.nT;
    return i * f.s_axis + t * f.s_t;
}

// order-preserving map double -> uint64 so that min / max can be done with integer atomics
__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

constexpr int RANGE_ROWS = 32;     // slice rows per CTA of the range pass

// R = max(max a, max clip(b)) - min(min a, min clip(b)) per slice (metrics.py:23-24).  Grid (slices, row
// chunks): a warp walks whole rows (lane = column, no index division), the CTA folds its result into the
// slice's {min key, max key} with integer atomics.  keys start at all-ones / zero.
template <class T>
__global__ void __launch_bounds__(256) ssim_range_kernel(const T* __restrict__ a, const T* __restrict__ b, SliceFamily f,
                                                          unsigned long long* __restrict__ keys) {
    __shared__ double scratch[32];
    const int64_t s = blockIdx.x;
    const int64_t base = slice_base(f, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t h0 = (int64_t)blockIdx.y * RANGE_ROWS;
    const int64_t h1 = h0 + RANGE_ROWS < f.H ? h0 + RANGE_ROWS : f.H;
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t h = h0 + warp; h < h1; h += 8) {
        const int64_t row = base + h * f.sh;
        for (int64_t w = lane; w < f.W; w += 32) {
            const double x = (double)a[row + w * f.sw];
            const double y = fmax((double)b[row + w * f.sw], 0.0);
            lo = fmin(lo, fmin(x, y));
            hi = fmax(hi, fmax(x, y));
        }
    }
    lo = block_min(lo, scratch);
    hi = block_max(hi, scratch);
    if (threadIdx.x == 0 && lo <= hi) {
        atomicMin(keys + 2 * s, dkey(lo));
        atomicMax(keys + 2 * s + 1, dkey(hi));
    }
}

// Horizontal pass: direct win-tap sums.  Vertical pass: a warp owns a strip of TILE_H / 8 output rows
// (lane = column) and slides the window down with one add and one subtract per quantity instead of
// `win` adds (for float32 inputs every intermediate is exact in float64).
template <class T>
__global__ void __launch_bounds__(256)
ssim_tile_kernel(const T* __restrict__ a, const T* __restrict__ b, SliceFamily f, const double* __restrict__ range,
                 int tiles_y, int tiles_x, double* __restrict__ partial) {
    extern __shared__ unsigned char tile_smem[];
    constexpr int HS_LD = TILE_W + 1;                                   // row stride of the horizontal sums (bank spread)
    double* hs = reinterpret_cast<double*>(tile_smem);                  // [5][PATCH_H][HS_LD]
    T* pa = reinterpret_cast<T*>(hs + 5 * PATCH_H * HS_LD);             // [PATCH_H][PATCH_W + 1]
    T* pb = pa + PATCH_H * (PATCH_W + 1);
    __shared__ double scratch[32];
    const int win = f.win, pad = (win - 1) / 2;
    int64_t bid = blockIdx.x;
    const int tx = (int)(bid % tiles_x); bid /= tiles_x;
    const int ty = (int)(bid % tiles_y); bid /= tiles_y;
    const int64_t s = bid;
    const int64_t base = slice_base(f, s);
    const int64_t oy = (int64_t)ty * TILE_H, ox = (int64_t)tx * TILE_W;   // interior coordinates
    const int ph = TILE_H + win - 1, pw = TILE_W + win - 1;
    for (int e = threadIdx.x; e < ph * pw; e += blockDim.x) {
        const int r = e / pw, c = e - r * pw;
        const int64_t h = oy + r, w = ox + c;
        T va = (T)0, vb = (T)0;
        if (h < f.H && w < f.W) {
            const int64_t off = base + h * f.sh + w * f.sw;
            va = a[off];
            vb = b[off];
            vb = vb > (T)0 ? vb : (T)0;
        }
        pa[r * (PATCH_W + 1) + c] = va;
        pb[r * (PATCH_W + 1) + c] = vb;
    }
    __syncthreads();
    // horizontal box sums: direct win-tap sums, every thread busy (a sliding sum along the row was measured
    // slower here: its dependent add chain serialises what this form does with pure instruction-level parallelism)
    for (int e = threadIdx.x; e < ph * TILE_W; e += blockDim.x) {
        const int r = e / TILE_W, c = e - r * TILE_W;
        const T* xa = pa + r * (PATCH_W + 1) + c;
        const T* xb = pb + r * (PATCH_W + 1) + c;
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        for (int j = 0; j < win; j++) {
            const double x = (double)xa[j], y = (double)xb[j];
            sx += x; sy += y;
            sxx = fma(x, x, sxx); syy = fma(y, y, syy); sxy = fma(x, y, sxy);
        }
        hs[(0 * PATCH_H + r) * HS_LD + c] = sx; hs[(1 * PATCH_H + r) * HS_LD + c] = sy; hs[(2 * PATCH_H + r) * HS_LD + c] = sxx;
        hs[(3 * PATCH_H + r) * HS_LD + c] = syy; hs[(4 * PATCH_H + r) * HS_LD + c] = sxy;
    }
    __syncthreads();
    const double R = range[s];
    const double c1 = (0.01 * R) * (0.01 * R), c2 = (0.03 * R) * (0.03 * R);
    const double np = (double)(win * win);
    const double inv_np = 1.0 / np, cov_norm = np / (np - 1.0);
    const int64_t ih = f.H - 2 * pad, iw = f.W - 2 * pad;   // interior extent
    double acc = 0.0;
    {
        constexpr int STRIP = TILE_H / 8;
        const int x = threadIdx.x & 31, y0 = (threadIdx.x >> 5) * STRIP;
        double v[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            double t = 0.0;
            for (int i = 0; i < win; i++) t += hs[(q * PATCH_H + y0 + i) * HS_LD + x];
            v[q] = t;
        }
#pragma unroll
        for (int dy = 0; dy < STRIP; dy++) {
            const int y = y0 + dy;
            if (dy > 0) {
#pragma unroll
                for (int q = 0; q < 5; q++)
                    v[q] += hs[(q * PATCH_H + y + win - 1) * HS_LD + x] - hs[(q * PATCH_H + y - 1) * HS_LD + x];
            }
            if (oy + y < ih && ox + x < iw) {
                const double ux = v[0] * inv_np, uy = v[1] * inv_np;
                const double vx = cov_norm * (v[2] * inv_np - ux * ux), vy = cov_norm * (v[3] * inv_np - uy * uy);
                const double vxy = cov_norm * (v[4] * inv_np - ux * uy);
                const double num = (2.0 * ux * uy + c1) * (2.0 * vxy + c2);
                const double den = (ux * ux + uy * uy + c1) * (vx + vy + c2);
                acc += num / den;
            }
        }
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

template <class T>
static int launch_ssim_tile(ndmps_ctx* ctx, const T* a, const T* b, const SliceFamily& f, const double* range, int ty, int tx,
                            int64_t nt, double* partial) {
    const size_t smem = (size_t)5 * PATCH_H * (TILE_W + 1) * sizeof(double) + (size_t)2 * PATCH_H * (PATCH_W + 1) * sizeof(T);
    NDMPS_TRY(raise_dynamic_smem((const void*)ssim_tile_kernel<T>, ctx->device, 96 * 1024));
    ssim_tile_kernel<T><<<(unsigned)nt, 256, smem, ctx->stream>>>(a, b, f, range, ty, tx, partial);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}

// ---------------------------------------------------------------------------------------------
// Slice-batched variants for families whose SLICE INDEX is the contiguous memory dimension
// (3-D: slices along the last axis; 4-D: every family, the frame index runs fastest).  A CTA
// takes 32 consecutive slices, lane = slice: every global load is a coalesced 128-byte line
// and the shared-memory patch [position][slice] is bank-conflict free.
// ---------------------------------------------------------------------------------------------
constexpr int BT_H = 8, BT_W = 8, BT_SLICES = 32;
constexpr int BT_RANGE_ROWS = 8;   // one slice row per warp in the batched range pass (few slice groups: many row chunks)


// keys[2s] = min key, keys[2s+1] = max key (initialised to all-ones / zero by the host).  lane = slice
// (32 consecutive slices are 32 consecutive addresses), a warp walks whole slice rows.
template <class T>
__global__ void __launch_bounds__(256) ssim_range_batched_kernel(const T* __restrict__ a, const T* __restrict__ b, SliceFamily f,
                                                                  unsigned long long* __restrict__ keys) {
    __shared__ double s_lo[8][BT_SLICES], s_hi[8][BT_SLICES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s = (int64_t)blockIdx.x * BT_SLICES + lane;
    double lo = INFINITY, hi = -INFINITY;
    if (s < f.S) {
        const int64_t base = slice_base(f, s);
        const int64_t h0 = (int64_t)blockIdx.y * BT_RANGE_ROWS;
        const int64_t h1 = h0 + BT_RANGE_ROWS < f.H ? h0 + BT_RANGE_ROWS : f.H;
        for (int64_t h = h0 + warp; h < h1; h += 8) {
            const int64_t row = base + h * f.sh;
#pragma unroll 8
            for (int64_t w = 0; w < f.W; w++) {
                const double x = (double)a[row + w * f.sw];
                const double y = fmax((double)b[row + w * f.sw], 0.0);
                lo = fmin(lo, fmin(x, y));
                hi = fmax(hi, fmax(x, y));
            }
        }
    }
    s_lo[warp][lane] = lo;
    s_hi[warp][lane] = hi;
    __syncthreads();
    if (warp == 0 && s < f.S) {
        for (int w = 1; w < 8; w++) { lo = fmin(lo, s_lo[w][lane]); hi = fmax(hi, s_hi[w][lane]); }
        if (lo <= hi) {
            atomicMin(keys + 2 * s, dkey(lo));
            atomicMax(keys + 2 * s + 1, dkey(hi));
        }
    }
}

__global__ void __launch_bounds__(256) ssim_range_finalize_kernel(const unsigned long long* __restrict__ keys, int64_t S,
                                                                   double* __restrict__ range) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) range[s] = dkey_inv(keys[2 * s + 1]) - dkey_inv(keys[2 * s]);
}

template <class T, int WIN>
__global__ void __launch_bounds__(256)
ssim_tile_batched_kernel(const T* __restrict__ a, const T* __restrict__ b, SliceFamily f, const double* __restrict__ range,
                         int tiles_y, int tiles_x, double* __restrict__ partial) {
    extern __shared__ unsigned char bt_smem[];
    constexpr int win = WIN, pad = (WIN - 1) / 2;
    const int ph = BT_H + win - 1, pw = BT_W + win - 1;
    T* pa = reinterpret_cast<T*>(bt_smem);                       // [ph*pw][32]
    T* pb = pa + (size_t)ph * pw * BT_SLICES;
    __shared__ double scratch[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t bid = blockIdx.x;
    const int tx = (int)(bid % tiles_x); bid /= tiles_x;
    const int ty = (int)(bid % tiles_y); bid /= tiles_y;
    const int64_t s = bid * BT_SLICES + lane;
    const bool live = s < f.S;
    const int64_t base = live ? slice_base(f, s) : 0;
    const int64_t oy = (int64_t)ty * BT_H, ox = (int64_t)tx * BT_W;
    for (int e = warp; e < ph * pw; e += 8) {
        const int r = e / pw, c = e - r * pw;
        const int64_t h = oy + r, w = ox + c;
        T va = (T)0, vb = (T)0;
        if (live && h < f.H && w < f.W) {
            const int64_t off = base + h * f.sh + w * f.sw;
            va = a[off];
            vb = b[off];
            vb = vb > (T)0 ? vb : (T)0;
        }
        pa[(size_t)e * BT_SLICES + lane] = va;
        pb[(size_t)e * BT_SLICES + lane] = vb;
    }
    __syncthreads();
    double acc = 0.0;
    if (live) {
        const double R = range[s];
        const double c1 = (0.01 * R) * (0.01 * R), c2 = (0.03 * R) * (0.03 * R);
        constexpr double np = (double)(WIN * WIN);
        constexpr double inv_np = 1.0 / np, cov_norm = np / (np - 1.0);
        const int64_t ih = f.H - 2 * pad, iw = f.W - 2 * pad;
        // one output column per warp: horizontal WIN-tap sums per patch row, vertical running sum over
        // a register ring of the last WIN rows (everything unrolled, the ring is static registers)
        const int x = warp;
        if (ox + x < iw) {
            double ring[WIN][5];
            double vs[5] = {0, 0, 0, 0, 0};
#pragma unroll
            for (int r = 0; r < BT_H + WIN - 1; r++) {
                double hsum[5] = {0, 0, 0, 0, 0};
                const size_t rowbase = ((size_t)r * pw + x) * BT_SLICES + lane;
#pragma unroll
                for (int j = 0; j < WIN; j++) {
                    const double xv = (double)pa[rowbase + (size_t)j * BT_SLICES];
                    const double yv = (double)pb[rowbase + (size_t)j * BT_SLICES];
                    hsum[0] += xv; hsum[1] += yv;
                    hsum[2] = fma(xv, xv, hsum[2]); hsum[3] = fma(yv, yv, hsum[3]); hsum[4] = fma(xv, yv, hsum[4]);
                }
#pragma unroll
                for (int q = 0; q < 5; q++) {
                    if (r >= WIN) vs[q] -= ring[r % WIN][q];
                    vs[q] += hsum[q];
                    ring[r % WIN][q] = hsum[q];
                }
                if (r >= WIN - 1) {
                    const int y = r - (WIN - 1);
                    if (oy + y < ih) {
                        const double ux = vs[0] * inv_np, uy = vs[1] * inv_np;
                        const double vx = cov_norm * (vs[2] * inv_np - ux * ux), vy = cov_norm * (vs[3] * inv_np - uy * uy);
                        const double vxy = cov_norm * (vs[4] * inv_np - ux * uy);
                        acc += ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
                    }
                }
            }
        }
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}


// ---------------------------------------------------------------------------------------------
// Streaming kernels for float32 inputs, window 7
// ---------------------------------------------------------------------------------------------
namespace stream {

constexpr int WIN = 7, PAD = 3, NP = 49;
constexpr int STRIP = 128 - (WIN - 1);         // output columns per warp of the column kernel (four columns per lane)
constexpr int BAND = 32;                        // output rows per unit (more, smaller units: the float64 pipe is the bound, whole waves matter)

// SSIM of one window from its float64 sums (x, y, xx + yy, xy); inv_r = 1 / data range of the slice
__device__ __forceinline__ float window_score(double sx, double sy, double sq, double sxy, double inv_r) {
    constexpr double inv_np = 1.0 / NP;
    constexpr float cov_norm = (float)NP / (NP - 1);
    const double ux = sx * inv_np, uy = sy * inv_np;
    const double vsum = fma(-uy, uy, fma(-ux, ux, sq * inv_np));           // (vx + vy) / cov_norm: the cancellation happens in float64
    const double vxy = fma(-ux, uy, sxy * inv_np);
    // range-normalised moments: C1 = 1e-4, C2 = 9e-4 are constants; everything after the cancellation in float32
    const float fr = (float)inv_r, fr2 = cov_norm * fr * fr;
    const float mx = (float)ux * fr, my = (float)uy * fr;
    const float vs = (float)vsum * fr2, vc = (float)vxy * fr2;
    const float num = (2.f * mx * my + 1e-4f) * (2.f * vc + 9e-4f);
    const float den = (mx * mx + my * my + 1e-4f) * (vs + 9e-4f);
    return num / den;
}

// lane = four consecutive columns (f.sw == 1).  Unit = (slice, strip of STRIP output columns, band of BAND output rows),
// one warp each: a row of the strip is one 16-byte load per lane and array.  Per row the column sums of the 7-row window
// slide (the row leaving the window is read again - it is still in L1/L2 - so no register ring is needed), then the
// seven-column window sums of a thread's four outputs come from its own four column sums plus six from the next two
// lanes (24 shuffles per quantity set instead of 4 per output).  partial[unit] = sum of the unit's window scores.
struct Quad { double q[4]; };     // x, y, xx + yy, xy of one column

__device__ __forceinline__ void quad_add(Quad& v, float xf, float yf, double sign) {
    const double x = (double)xf, y = (double)fmaxf(yf, 0.f);
    v.q[0] = fma(sign, x, v.q[0]);
    v.q[1] = fma(sign, y, v.q[1]);
    v.q[2] = fma(sign, fma(x, x, y * y), v.q[2]);
    v.q[3] = fma(sign * x, y, v.q[3]);
}

__device__ __forceinline__ float4 load4(const float* p, int64_t col, int64_t W, bool vec) {
    if (vec && col + 4 <= W) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < W) v.x = __ldg(p);
    if (col + 1 < W) v.y = __ldg(p + 1);
    if (col + 2 < W) v.z = __ldg(p + 2);
    if (col + 3 < W) v.w = __ldg(p + 3);
    return v;
}

__global__ void __launch_bounds__(256, 2)
ssim_col_kernel(const float* __restrict__ a, const float* __restrict__ b, SliceFamily f, const double* __restrict__ range,
                int n_strips, int n_bands, int64_t n_units, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (u >= n_units) return;
    const int rb = (int)(u % n_bands);
    const int cs = (int)((u / n_bands) % n_strips);
    const int64_t s = u / ((int64_t)n_bands * n_strips);
    const int64_t base = slice_base(f, s);
    const int64_t ih = f.H - 2 * PAD, iw = f.W - 2 * PAD;
    const int64_t oy0 = (int64_t)rb * BAND, oy1 = oy0 + BAND < ih ? oy0 + BAND : ih;     // output rows [oy0, oy1)
    const int64_t col = (int64_t)cs * STRIP + 4 * lane;                                   // first of this lane's four input columns
    const double inv_r = 1.0 / range[s];
    const float* pa = a + base + col;
    const float* pb = b + base + col;
    // 16-byte loads need 16-byte aligned rows
    const bool vec = ((f.sh & 3) == 0) && (((base + col) & 3) == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    Quad v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k].q[0] = v[k].q[1] = v[k].q[2] = v[k].q[3] = 0.0;
    double acc = 0.0;
    const int64_t nrows = oy1 - oy0 + (WIN - 1);
    float4 xa = load4(pa + oy0 * f.sh, col, f.W, vec), xb = load4(pb + oy0 * f.sh, col, f.W, vec);
    for (int64_t r = 0; r < nrows; r++) {
        // prefetch the next incoming row and this row's outgoing one before consuming the current
        float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na, oa = na, ob = na;
        if (r + 1 < nrows) { na = load4(pa + (oy0 + r + 1) * f.sh, col, f.W, vec); nb = load4(pb + (oy0 + r + 1) * f.sh, col, f.W, vec); }
        const bool drop = r >= WIN;
        if (drop) { oa = load4(pa + (oy0 + r - WIN) * f.sh, col, f.W, vec); ob = load4(pb + (oy0 + r - WIN) * f.sh, col, f.W, vec); }
        quad_add(v[0], xa.x, xb.x, 1.0); quad_add(v[1], xa.y, xb.y, 1.0); quad_add(v[2], xa.z, xb.z, 1.0); quad_add(v[3], xa.w, xb.w, 1.0);
        if (drop) {
            quad_add(v[0], oa.x, ob.x, -1.0); quad_add(v[1], oa.y, ob.y, -1.0); quad_add(v[2], oa.z, ob.z, -1.0); quad_add(v[3], oa.w, ob.w, -1.0);
        }
        if (r >= WIN - 1) {
            double hs[4][4];                                  // [output column][quantity]
#pragma unroll
            for (int q = 0; q < 4; q++) {
                // columns 4..7 from the next lane, 8..9 from the one after
                const double c4 = __shfl_down_sync(0xffffffffu, v[0].q[q], 1), c5 = __shfl_down_sync(0xffffffffu, v[1].q[q], 1);
                const double c6 = __shfl_down_sync(0xffffffffu, v[2].q[q], 1), c7 = __shfl_down_sync(0xffffffffu, v[3].q[q], 1);
                const double c8 = __shfl_down_sync(0xffffffffu, v[0].q[q], 2), c9 = __shfl_down_sync(0xffffffffu, v[1].q[q], 2);
                const double h0 = ((v[0].q[q] + v[1].q[q]) + (v[2].q[q] + v[3].q[q])) + ((c4 + c5) + c6);
                const double h1 = h0 - v[0].q[q] + c7, h2 = h1 - v[1].q[q] + c8, h3 = h2 - v[2].q[q] + c9;
                hs[0][q] = h0; hs[1][q] = h1; hs[2][q] = h2; hs[3][q] = h3;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t oc = col + k;                   // output column (window starts here)
                if (4 * lane + k < STRIP && oc < iw) acc += (double)window_score(hs[k][0], hs[k][1], hs[k][2], hs[k][3], inv_r);
            }
        }
        xa = na; xb = nb;
    }
    acc = warp_sum(acc);
    if (lane == 0) partial[u] = acc;
}

// lane = slice (32 consecutive slices are 32 consecutive addresses).  Unit = (group of 32 slices, output column,
// band of BAND output rows), one warp each; every thread owns one slice's column: seven loads per array and row.
// per_slice: partial[slice * units_per_slice + unit_in_slice] = sum of the window scores (for ssim_slices);
// otherwise partial[unit] = the same summed over the unit's 32 slices.
__global__ void __launch_bounds__(256, 2)
ssim_slice_kernel_f32(const float* __restrict__ a, const float* __restrict__ b, SliceFamily f, const double* __restrict__ range,
                      int n_cols, int n_bands, int64_t n_units, int per_slice, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (u >= n_units) return;
    const int rb = (int)(u % n_bands);
    const int oc = (int)((u / n_bands) % n_cols);                                         // output column
    const int64_t g = u / ((int64_t)n_bands * n_cols);
    const int64_t s = g * 32 + lane;
    const bool live = s < f.S;                             // dead lanes of the last slice group only join the final reduction
    const int64_t base = live ? slice_base(f, s) : 0;
    const int64_t ih = f.H - 2 * PAD;
    const int64_t oy0 = (int64_t)rb * BAND, oy1 = oy0 + BAND < ih ? oy0 + BAND : ih;
    const double inv_r = live ? 1.0 / range[s] : 0.0;
    const float* pa = a + base + (int64_t)oc * f.sw;
    const float* pb = b + base + (int64_t)oc * f.sw;
    double ring[WIN][4];
#pragma unroll
    for (int j = 0; j < WIN; j++) ring[j][0] = ring[j][1] = ring[j][2] = ring[j][3] = 0.0;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    double acc = 0.0;
    const int64_t nrows = live ? oy1 - oy0 + (WIN - 1) : 0;
    for (int64_t r0 = 0; r0 < nrows; r0 += WIN) {
#pragma unroll
        for (int j = 0; j < WIN; j++) {
            const int64_t r = r0 + j;
            if (r < nrows) {
                const int64_t off = (oy0 + r) * f.sh;
                float xf[WIN], yf[WIN];
#pragma unroll
                for (int t = 0; t < WIN; t++) { xf[t] = __ldg(pa + off + t * f.sw); yf[t] = __ldg(pb + off + t * f.sw); }
                asm volatile("" ::: "memory");           // keep the loads of later rows from being hoisted above (register pressure)
                double h0 = 0.0, h1 = 0.0, h2 = 0.0, h3 = 0.0;
#pragma unroll
                for (int t = 0; t < WIN; t++) {
                    const double x = (double)xf[t], y = (double)fmaxf(yf[t], 0.f);
                    h0 += x; h1 += y; h2 += fma(x, x, y * y); h3 = fma(x, y, h3);
                }
                v0 += h0 - ring[j][0]; v1 += h1 - ring[j][1]; v2 += h2 - ring[j][2]; v3 += h3 - ring[j][3];
                ring[j][0] = h0; ring[j][1] = h1; ring[j][2] = h2; ring[j][3] = h3;
                if (r >= WIN - 1) acc += (double)window_score(v0, v1, v2, v3, inv_r);
            }
        }
    }
    if (per_slice) {
        if (live) partial[s * ((int64_t)n_cols * n_bands) + (int64_t)oc * n_bands + rb] = acc;
    } else {                                               // family total: the 32 slices of the unit summed in lane order
        acc = warp_sum(acc);
        if (lane == 0) partial[u] = acc;
    }
}

}  // namespace stream

// out[fam] = sum(partial[begin..end)) * scale, one CTA per family
struct FinalArgs { int64_t bounds[4]; double scale[3]; };   // by value: no host -> device copy behind other streams' uploads
__global__ void __launch_bounds__(256) ssim_final_kernel(const double* __restrict__ partial, const FinalArgs fa, double* __restrict__ out) {
    __shared__ double scratch[32];
    const int fam = blockIdx.x;
    double acc = 0.0;
    for (int64_t i = fa.bounds[fam] + threadIdx.x; i < fa.bounds[fam + 1]; i += blockDim.x) acc += partial[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[fam] = acc * fa.scale[fam];
}

// scores[s] = sum of the slice's tile partials * scale, one CTA per slice
__global__ void __launch_bounds__(128) ssim_slice_kernel(const double* __restrict__ partial, int tiles_per_slice, double scale,
                                                          double* __restrict__ scores) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < tiles_per_slice; i += blockDim.x) acc += partial[(int64_t)blockIdx.x * tiles_per_slice + i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) scores[blockIdx.x] = acc * scale;
}

// per-slice data range of a family into range[0..S): keys initialised, range pass, finalize
template <class T>
static int slice_ranges(ndmps_ctx* ctx, const T* a, const T* b, const SliceFamily& f, bool batched, double* range) {
    unsigned long long* keys = nullptr;
    NDMPS_TRY(ctx->ws.get<unsigned long long>((size_t)(2 * f.S), &keys));
    // min keys start at all-ones, max keys at zero: 0xFF.. / 0x00.. interleaved per slice
    NDMPS_CUDA_TRY(cudaMemset2DAsync(keys, 16, 0xFF, 8, (size_t)f.S, ctx->stream));
    NDMPS_CUDA_TRY(cudaMemset2DAsync(keys + 1, 16, 0x00, 8, (size_t)f.S, ctx->stream));
    const unsigned chunks = (unsigned)((f.H + RANGE_ROWS - 1) / RANGE_ROWS);
    if (batched) {
        dim3 rgrid((unsigned)((f.S + BT_SLICES - 1) / BT_SLICES), (unsigned)((f.H + BT_RANGE_ROWS - 1) / BT_RANGE_ROWS));
        ssim_range_batched_kernel<T><<<rgrid, 256, 0, ctx->stream>>>(a, b, f, keys);
    } else {
        dim3 rgrid((unsigned)f.S, chunks);
        ssim_range_kernel<T><<<rgrid, 256, 0, ctx->stream>>>(a, b, f, keys);
    }
    NDMPS_LAUNCH_CHECK(ctx);
    ssim_range_finalize_kernel<<<(unsigned)((f.S + 255) / 256), 256, 0, ctx->stream>>>(keys, f.S, range);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}

// ---- streaming path: which kernel serves a family, how many partials it writes, launch ------------------
static inline bool slices_contiguous(const SliceFamily& f) { return (f.nT > 1 && f.s_t == 1) || (f.nT == 1 && f.s_axis == 1); }
static inline int stream_kind(const ndmps_ctx* ctx, const SliceFamily& f, const float*) {
    if (ctx->opt_ssim_exact || f.win != 7) return 0;
    if (f.sw == 1) return 1;                          // lane = column
    if (slices_contiguous(f)) return 2;               // lane = slice
    return 0;
}
static inline int stream_kind(const ndmps_ctx*, const SliceFamily&, const double*) { return 0; }

struct StreamPlan { int kind = 0, n_a = 0, n_bands = 0, by_slice = 0; int64_t n_units = 0, n_partials = 0, per_slice = 0; };

static StreamPlan stream_plan(int kind, const SliceFamily& f, bool by_slice) {
    StreamPlan p;
    p.kind = kind;
    p.by_slice = by_slice ? 1 : 0;
    const int64_t ih = f.H - 6, iw = f.W - 6;
    p.n_bands = (int)((ih + stream::BAND - 1) / stream::BAND);
    if (kind == 1) {
        p.n_a = (int)((iw + stream::STRIP - 1) / stream::STRIP);                 // strips
        p.per_slice = (int64_t)p.n_a * p.n_bands;
        p.n_units = f.S * p.per_slice;
        p.n_partials = p.n_units;
    } else {
        p.n_a = (int)iw;                                                          // output columns
        p.per_slice = (int64_t)p.n_a * p.n_bands;
        p.n_units = ((f.S + 31) / 32) * p.per_slice;
        p.n_partials = by_slice ? f.S * p.per_slice : p.n_units;
    }
    return p;
}

static int launch_stream(ndmps_ctx* ctx, const float* a, const float* b, const SliceFamily& f, const StreamPlan& p, const double* range,
                         double* partial) {
    const unsigned grid = (unsigned)((p.n_units + 7) / 8);
    if (p.kind == 1) stream::ssim_col_kernel<<<grid, 256, 0, ctx->stream>>>(a, b, f, range, p.n_a, p.n_bands, p.n_units, partial);
    else stream::ssim_slice_kernel_f32<<<grid, 256, 0, ctx->stream>>>(a, b, f, range, p.n_a, p.n_bands, p.n_units, p.by_slice, partial);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}
static int launch_stream(ndmps_ctx*, const double*, const double*, const SliceFamily&, const StreamPlan&, const double*, double*) {
    return NDMPS_ERR_INVALID;                          // never selected for float64 inputs
}

template <class T>
static int ssim_slices_typed(ndmps_ctx* ctx, const T* a, const T* b, const SliceFamily& f, double* scores_host) {
    NDMPS_TRY(ensure_pinned(ctx, (size_t)f.S + 64));
    int pad = (f.win - 1) / 2;
    int64_t ih = f.H - 2 * pad, iw = f.W - 2 * pad;
    if (const int kind = stream_kind(ctx, f, a)) {
        const StreamPlan p = stream_plan(kind, f, true);
        NDMPS_REQUIRE(p.n_units < (int64_t(1) << 33) && p.per_slice < (int64_t(1) << 31), "ssim: too many units");
        double *range = nullptr, *partial = nullptr, *scores = nullptr;
        NDMPS_TRY(ctx->ws.get<double>((size_t)f.S, &range));
        NDMPS_TRY(ctx->ws.get<double>((size_t)p.n_partials, &partial));
        NDMPS_TRY(ctx->ws.get<double>((size_t)f.S, &scores));
        NDMPS_TRY(slice_ranges<T>(ctx, a, b, f, kind == 2, range));
        NDMPS_TRY(launch_stream(ctx, a, b, f, p, range, partial));
        ssim_slice_kernel<<<(unsigned)f.S, 128, 0, ctx->stream>>>(partial, (int)p.per_slice, 1.0 / ((double)ih * (double)iw), scores);
        NDMPS_LAUNCH_CHECK(ctx);
        NDMPS_TRY(readback(ctx, ctx->pinned, scores, (size_t)f.S * sizeof(double)));
        NDMPS_CUDA_TRY(stream_wait(ctx));
        memcpy(scores_host, ctx->pinned, (size_t)f.S * sizeof(double));
        return NDMPS_OK;
    }
    int ty = (int)((ih + TILE_H - 1) / TILE_H), tx = (int)((iw + TILE_W - 1) / TILE_W);
    int64_t nt = f.S * ty * tx;
    NDMPS_REQUIRE(nt < (int64_t(1) << 31), "ssim: too many tiles");
    double *range = nullptr, *partial = nullptr, *scores = nullptr;
    NDMPS_TRY(ctx->ws.get<double>((size_t)f.S, &range));
    NDMPS_TRY(ctx->ws.get<double>((size_t)nt, &partial));
    NDMPS_TRY(ctx->ws.get<double>((size_t)f.S, &scores));
    NDMPS_REQUIRE(f.S < 65536 * 32, "ssim: too many slices");
    NDMPS_TRY(slice_ranges<T>(ctx, a, b, f, false, range));
    NDMPS_TRY(launch_ssim_tile<T>(ctx, a, b, f, range, ty, tx, nt, partial));
    ssim_slice_kernel<<<(unsigned)f.S, 128, 0, ctx->stream>>>(partial, ty * tx, 1.0 / ((double)ih * (double)iw), scores);
    NDMPS_LAUNCH_CHECK(ctx);
    NDMPS_TRY(readback(ctx, ctx->pinned, scores, (size_t)f.S * sizeof(double)));
    NDMPS_CUDA_TRY(stream_wait(ctx));
    memcpy(scores_host, ctx->pinned, (size_t)f.S * sizeof(double));
    return NDMPS_OK;
}

template <class T>
static int ssim_typed(ndmps_ctx* ctx, const T* a, const T* b, int nfam, const SliceFamily* fams, double* out_host) {
    NDMPS_TRY(ensure_pinned(ctx, 64));
    int64_t bounds_h[4] = {0, 0, 0, 0};
    double scale_h[3] = {0, 0, 0};
    int tiles_y[3], tiles_x[3];
    bool batched[3];
    StreamPlan splan[3];
    int64_t total_tiles = 0, total_slices = 0;
    for (int k = 0; k < nfam; k++) {
        const SliceFamily& f = fams[k];
        int pad = (f.win - 1) / 2;
        int64_t ih = f.H - 2 * pad, iw = f.W - 2 * pad;
        if (const int kind = stream_kind(ctx, f, a)) {      // float32, window 7: streaming kernel, partials = its units
            splan[k] = stream_plan(kind, f, false);
            batched[k] = kind == 2;
            tiles_y[k] = tiles_x[k] = 0;
            bounds_h[k] = total_tiles;
            total_tiles += splan[k].n_partials;
            bounds_h[k + 1] = total_tiles;
            scale_h[k] = 1.0 / ((double)f.S * (double)ih * (double)iw);
            total_slices += f.S;
            continue;
        }
        // consecutive slice indices are consecutive addresses: frames (4-D) or the last axis (3-D)
        batched[k] = f.win == 7 && f.S >= 8 && ((f.nT > 1 && f.s_t == 1) || (f.nT == 1 && f.s_axis == 1));
        const int th = batched[k] ? BT_H : TILE_H, tw = batched[k] ? BT_W : TILE_W;
        tiles_y[k] = (int)((ih + th - 1) / th);
        tiles_x[k] = (int)((iw + tw - 1) / tw);
        int64_t nt = (batched[k] ? (f.S + BT_SLICES - 1) / BT_SLICES : f.S) * tiles_y[k] * tiles_x[k];
        NDMPS_REQUIRE(nt < (int64_t(1) << 31), "ssim: too many tiles");
        bounds_h[k] = total_tiles;
        total_tiles += nt;
        bounds_h[k + 1] = total_tiles;
        scale_h[k] = 1.0 / ((double)f.S * (double)ih * (double)iw);
        total_slices += f.S;
    }
    double *range = nullptr, *partial = nullptr, *out_dev = nullptr;
    NDMPS_TRY(ctx->ws.get<double>((size_t)total_slices, &range));
    NDMPS_TRY(ctx->ws.get<double>((size_t)total_tiles, &partial));
    NDMPS_TRY(ctx->ws.get<double>(4, &out_dev));
    FinalArgs fa;
    for (int k = 0; k < 4; k++) fa.bounds[k] = bounds_h[k];
    for (int k = 0; k < 3; k++) fa.scale[k] = scale_h[k];
    int64_t slice_off = 0;
    for (int k = 0; k < nfam; k++) {
        const SliceFamily& f = fams[k];
        int64_t nt = bounds_h[k + 1] - bounds_h[k];
        if (splan[k].kind) {
            NDMPS_TRY(slice_ranges<T>(ctx, a, b, f, splan[k].kind == 2, range + slice_off));
            NDMPS_TRY(launch_stream(ctx, a, b, f, splan[k], range + slice_off, partial + bounds_h[k]));
        } else if (batched[k]) {
            const size_t bsm = (size_t)2 * (BT_H + f.win - 1) * (BT_W + f.win - 1) * BT_SLICES * sizeof(T);
            NDMPS_TRY(raise_dynamic_smem((const void*)ssim_tile_batched_kernel<T, 7>, ctx->device, 112 * 1024));
            NDMPS_TRY(slice_ranges<T>(ctx, a, b, f, true, range + slice_off));
            ssim_tile_batched_kernel<T, 7><<<(unsigned)nt, 256, bsm, ctx->stream>>>(a, b, f, range + slice_off, tiles_y[k], tiles_x[k],
                                                                                 partial + bounds_h[k]);
            NDMPS_LAUNCH_CHECK(ctx);
        } else {
            NDMPS_TRY(slice_ranges<T>(ctx, a, b, f, false, range + slice_off));
            NDMPS_TRY(launch_ssim_tile<T>(ctx, a, b, f, range + slice_off, tiles_y[k], tiles_x[k], nt, partial + bounds_h[k]));
        }
        slice_off += f.S;
    }
    ssim_final_kernel<<<nfam, 256, 0, ctx->stream>>>(partial, fa, out_dev);
    NDMPS_LAUNCH_CHECK(ctx);
    NDMPS_TRY(readback(ctx, ctx->pinned, out_dev, 4 * sizeof(double)));
    NDMPS_CUDA_TRY(stream_wait(ctx));
    double m = 0.0;
    for (int k = 0; k < nfam; k++) m += ctx->pinned[k];
    out_host[0] = m / nfam;
    return NDMPS_OK;
}

static int pick_win(int64_t h, int64_t w) {
    int64_t m = h < w ? h : w;
    int win = (int)(m < 7 ? m : 7);
    if (win % 2 == 0) win -= 1;
    return win;
}

}  // namespace ndmps

using namespace ndmps;

extern "C" {

int ndmps_ssim(ndmps_ctx_t* ctx, const void* a, const void* b, int dtype, int ndim, const int64_t* shape, double* out_host) {
    NDMPS_REQUIRE(ctx && a && b && shape && out_host, "ndmps_ssim: NULL argument");
    NDMPS_REQUIRE(dtype_ok(dtype), "ndmps_ssim: bad dtype");
    NDMPS_REQUIRE(ndim >= 2 && ndim <= 4, "Unsupported tensor dimension for SSIM: %d", ndim);
    for (int i = 0; i < ndim; i++) NDMPS_REQUIRE(shape[i] >= 1, "ndmps_ssim: non-positive extent");
    NDMPS_TRY(ctx->ws.reset(ctx->stream));
    SliceFamily fams[3];
    int nfam = 0;
    if (ndim == 2) {
        SliceFamily f{1, shape[0], shape[1], shape[1], 1, 1, 0, 0, pick_win(shape[0], shape[1])};
        fams[nfam++] = f;
    } else {
        const int64_t n0 = shape[0], n1 = shape[1], n2 = shape[2], nT = ndim == 4 ? shape[3] : 1;
        const int64_t s2 = nT, s1 = n2 * nT, s0 = n1 * n2 * nT;
        fams[nfam++] = SliceFamily{n0 * nT, n1, n2, s1, s2, nT, s0, 1, pick_win(n1, n2)};   // original[i]
        fams[nfam++] = SliceFamily{n1 * nT, n0, n2, s0, s2, nT, s1, 1, pick_win(n0, n2)};   // original[:, i, :]
        fams[nfam++] = SliceFamily{n2 * nT, n0, n1, s0, s1, nT, s2, 1, pick_win(n0, n1)};   // original[:, :, i]
    }
    for (int k = 0; k < nfam; k++) {
        if (fams[k].win < 3) {   // a 1x1 window has no sample covariance: the reference yields NaN
            out_host[0] = NAN;
            return NDMPS_OK;
        }
    }
    StageScope sc(ctx, ST_METRIC);
    if (dtype == NDMPS_F32) return ssim_typed<float>(ctx, (const float*)a, (const float*)b, nfam, fams, out_host);
    return ssim_typed<double>(ctx, (const double*)a, (const double*)b, nfam, fams, out_host);
}

int ndmps_ssim_slices(ndmps_ctx_t* ctx, const void* a, const void* b, int dtype, const int64_t* shape, int axis,
                      double* scores_out_host) {
    NDMPS_REQUIRE(ctx && a && b && shape && scores_out_host, "ndmps_ssim_slices: NULL argument");
    NDMPS_REQUIRE(dtype_ok(dtype), "ndmps_ssim_slices: bad dtype");
    NDMPS_REQUIRE(axis >= 0 && axis <= 2, "Invalid axis %d for 3D SSIM.", axis);
    for (int i = 0; i < 3; i++) NDMPS_REQUIRE(shape[i] >= 1, "ndmps_ssim_slices: non-positive extent");
    NDMPS_TRY(ctx->ws.reset(ctx->stream));
    const int64_t n0 = shape[0], n1 = shape[1], n2 = shape[2];
    const int64_t s2 = 1, s1 = n2, s0 = n1 * n2;
    SliceFamily f;
    if (axis == 0) f = SliceFamily{n0, n1, n2, s1, s2, 1, s0, 1, pick_win(n1, n2)};
    else if (axis == 1) f = SliceFamily{n1, n0, n2, s0, s2, 1, s1, 1, pick_win(n0, n2)};
    else f = SliceFamily{n2, n0, n1, s0, s1, 1, s2, 1, pick_win(n0, n1)};
    if (f.win < 3) {
        for (int64_t i = 0; i < f.S; i++) scores_out_host[i] = NAN;
        return NDMPS_OK;
    }
    if (dtype == NDMPS_F32) return ssim_slices_typed<float>(ctx, (const float*)a, (const float*)b, f, scores_out_host);
    return ssim_slices_typed<double>(ctx, (const double*)a, (const double*)b, f, scores_out_host);
}

}  // extern "C"
