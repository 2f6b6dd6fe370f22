// K3b: the leading k eigenpairs of a bond-sized Gram matrix, float64.
//
// When from_dense / compress run with a bond cap (core/ndmps.py:74 via quimb's max_bond; the
// reference benchmark's chi = 64 on (chi d) x (chi d) = 512 x 512 Gram matrices), only the k
// largest singular triplets of an unfolding are kept, and the trimming rule
// (quimb.tensor.decomp, Appendix A.1 of SURVEY.md) needs nothing else from the discarded part than its
// total weight, which is trace(G) minus the kept eigenvalues.  So instead of diagonalising
// the whole matrix (eig.cu: ~n sequential rounds per sweep, ~11 sweeps) this path does
//
//   1. G = Q T Q^T          Householder tridiagonalisation, ONE persistent cooperative launch,
//                           one grid barrier per column (tridiag_kernel);
//   2. k shifts             bisection on T, 129-section per pass (bisect_kernel);
//   3. k vectors of T       inverse iteration (pivoted tridiagonal LU, one CTA per shift),
//                           the block re-orthonormalised by Cholesky-QR between iterations so
//                           unresolved clusters keep spanning their invariant subspace;
//   4. Rayleigh-Ritz        H = Q_k^T T Q_k (k x k) through the Jacobi solver of eig.cu: the
//                           returned eigenvalues are second-order accurate Ritz values and the
//                           basis is rotated inside clusters;
//   5. U = Q Z              the Householder reflectors applied to the k Ritz vectors.
//
// n - 2 sequential steps instead of ~11 n, and everything after step 1 works on k vectors.
// The caller (ttsvd.cu) checks on the host that the cap really binds (discarded weight well
// above the cutoff target, lambda_k well above the rounding floor of T) and otherwise falls
// back to the full solver, so rank decisions stay identical to the oracle's.
#include "common.cuh"

namespace ndmps {

namespace topk {

constexpr int TDT = 256;   // threads per CTA of the tridiagonalisation
constexpr int KM = 4;      // vector elements per thread -> n <= 1024
constexpr int BIS = 128;   // shifts per pass and eigenvalue in the bisection

__device__ __forceinline__ double rcp_fast(double x) {      // MUFU seed + 2 Newton steps, normal range
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned*)counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// sum over the CTA's 8 warps, same value (bitwise) in every thread; one __syncthreads
__device__ __forceinline__ double block_sum_all(double v, double* scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < TDT / 32; w++) s += scratch[w];
    return s;
}

// ---------------------------------------------------------------------------------------------
// 1. Householder tridiagonalisation.  Rows are dealt cyclically to the CTAs (row i lives in CTA
// i mod C, shared memory, full rows: the symmetric half is not exploited, the 2x flops are
// free next to the barrier).  Per column jn every CTA redundantly derives, from the partial
// products p = A u and the pivot row that were published before the barrier, the vector w of
// the rank-2 update of the PREVIOUS reflector and the NEXT reflector u'; then one fused pass
// over its rows applies A -= u w^T + w u^T and accumulates p' = A u', publishes p' and the next
// pivot row, and meets the others at the barrier.  One barrier per column.
// V row j = reflector j (unit at j+1, zero before), H_j = I - tau_j v_j v_j^T.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TDT)
tridiag_kernel(const double* __restrict__ G, int n, int nr, double* __restrict__ V, double* __restrict__ tau_g,
               double* __restrict__ d_g, double* __restrict__ e_g, double* pbuf, double* rowbuf, unsigned* ctrl) {
    extern __shared__ double sm[];
    double* A = sm;                                      // nr x n
    double* ub0 = A + (size_t)nr * n;                    // n
    double* ub1 = ub0 + n;                               // n
    double* wv = ub1 + n;                                // n
    __shared__ double red0[TDT / 32], red1[TDT / 32];
    __shared__ double s_alpha;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x, C = gridDim.x;
    for (int rl = 0; rl < nr; rl++) {
        const int gi = rl * C + cta;
        if (gi >= n) break;
        for (int c = tid; c < n; c += TDT) A[(size_t)rl * n + c] = G[(size_t)gi * n + c];
    }
    for (int c = tid; c < n; c += TDT) { ub0[c] = 0.0; ub1[c] = 0.0; wv[c] = 0.0; }
    double* u = ub0;
    double* un = ub1;
    double tau = 0.0;
    unsigned epoch = 0;
    __syncthreads();
    for (int jn = 0; jn <= n - 2; jn++) {
        const int par = jn & 1;
        double r[KM];                                    // row jn of the current matrix, columns >= jn
        if (jn == 0) {
#pragma unroll
            for (int m = 0; m < KM; m++) {
                const int k = tid + m * TDT;
                r[m] = k < n ? G[k] : 0.0;
            }
        } else {
            const double* pb = pbuf + (size_t)(par ^ 1) * n;
            const double* rb = rowbuf + (size_t)(par ^ 1) * n;
            double pf[KM], rj[KM];
#pragma unroll
            for (int m = 0; m < KM; m++) {
                const int k = tid + m * TDT;
                const bool on = k >= jn && k < n;
                pf[m] = on ? __ldcg(pb + k) : 0.0;
                rj[m] = on ? __ldcg(rb + k) : 0.0;
            }
            const double pj = __ldcg(pb + jn);
            double part = 0.0;
#pragma unroll
            for (int m = 0; m < KM; m++) {
                const int k = tid + m * TDT;
                if (k < n) part = fma(pf[m], u[k], part);
            }
            const double s = block_sum_all(part, red0);
            const double K = -0.5 * tau * tau * s;
            const double wj = fma(tau, pj, K);           // u[jn] = 1
#pragma unroll
            for (int m = 0; m < KM; m++) {
                const int k = tid + m * TDT;
                r[m] = 0.0;
                if (k >= jn && k < n) {
                    const double uk = u[k];
                    const double wk = fma(tau, pf[m], K * uk);
                    wv[k] = wk;
                    r[m] = rj[m] - wk - wj * uk;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < KM; m++) {
            const int k = tid + m * TDT;
            if (k == jn && cta == 0) d_g[jn] = r[m];
            if (k == jn + 1) s_alpha = r[m];
        }
        if (jn == n - 2) {                               // last 2 x 2 block: no reflector left
            __syncthreads();
            if (cta == 0 && tid == 0) { e_g[jn] = s_alpha; tau_g[jn] = 0.0; }
            if ((n - 1) % C == cta && tid == 0) {
                double v = A[(size_t)((n - 1) / C) * n + (n - 1)];
                if (jn > 0) v -= 2.0 * u[n - 1] * wv[n - 1];
                d_g[n - 1] = v;
            }
            break;
        }
        double part = 0.0;
#pragma unroll
        for (int m = 0; m < KM; m++) {
            const int k = tid + m * TDT;
            if (k >= jn + 2 && k < n) part = fma(r[m], r[m], part);
        }
        const double sigma = block_sum_all(part, red1);  // the barrier inside also publishes s_alpha and wv
        const double alpha = s_alpha;
        double beta = alpha, taun = 0.0, scal = 0.0;
        if (sigma > 1e-280) {
            const double nrm = sqrt(fma(alpha, alpha, sigma));
            beta = -copysign(nrm, alpha);
            taun = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
#pragma unroll
        for (int m = 0; m < KM; m++) {
            const int k = tid + m * TDT;
            if (k < n) {
                const double v = k == jn + 1 ? 1.0 : (k >= jn + 2 ? r[m] * scal : 0.0);
                un[k] = v;
                if (cta == jn % C) V[(size_t)jn * n + k] = v;
            }
        }
        if (cta == 0 && tid == 0) { e_g[jn] = beta; tau_g[jn] = taun; }
        __syncthreads();
        // fused rank-2 update (reflector jn-1) + product with reflector jn, rows and columns >= jn+1
        double* pw = pbuf + (size_t)par * n;
        double* rw = rowbuf + (size_t)par * n;
        const int kbase = (jn + 1) & ~31;
        for (int rl = warp; rl < nr; rl += TDT / 32) {
            const int gi = rl * C + cta;
            if (gi >= n || gi < jn + 1) continue;
            double* Ar = A + (size_t)rl * n;
            const bool pub = gi == jn + 1;
            double acc = 0.0;
            if (jn > 0) {
                const double ui = u[gi], wi = wv[gi];
                for (int k = kbase + lane; k < n; k += 32) {
                    if (k >= jn + 1) {
                        double a = Ar[k];
                        a = fma(-ui, wv[k], a);
                        a = fma(-wi, u[k], a);
                        Ar[k] = a;
                        acc = fma(a, un[k], acc);
                        if (pub) __stcg(rw + k, a);
                    }
                }
            } else {
                for (int k = kbase + lane; k < n; k += 32) {
                    if (k >= jn + 1) {
                        const double a = Ar[k];
                        acc = fma(a, un[k], acc);
                        if (pub) __stcg(rw + k, a);
                    }
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) __stcg(pw + gi, acc);
        }
        epoch++;
        grid_barrier(ctrl, epoch * C);
        double* t = u; u = un; un = t;
        tau = taun;
    }
}

// ---------------------------------------------------------------------------------------------
// 2. Bisection: CTA t finds the t-th largest eigenvalue of T.  Every pass evaluates the Sturm
// count at 128 interior points of the bracket (one per thread), so a pass divides the bracket
// by 129.  bounds[0] = max |Gershgorin bound| (scale of T) for the later kernels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIS)
bisect_kernel(const double* __restrict__ d_g, const double* __restrict__ e_g, int n, int passes, double* __restrict__ lam,
              double* __restrict__ bounds) {
    extern __shared__ double sm[];
    double* d = sm;            // n
    double* e2 = sm + n;       // n (e2[n-1] = 0)
    __shared__ double scratch[32];
    __shared__ double s_lo, s_hi, s_e2max;
    const int tid = threadIdx.x;
    double gl = INFINITY, gu = -INFINITY, em = 0.0;
    for (int i = tid; i < n; i += BIS) {
        const double di = d_g[i];
        const double el = i > 0 ? fabs(e_g[i - 1]) : 0.0, er = i < n - 1 ? fabs(e_g[i]) : 0.0;
        d[i] = di;
        e2[i] = er * er;
        gl = fmin(gl, di - el - er);
        gu = fmax(gu, di + el + er);
        em = fmax(em, er * er);
    }
    gl = block_min(gl, scratch);
    if (tid == 0) s_lo = gl;
    gu = block_max(gu, scratch);
    if (tid == 0) s_hi = gu;
    em = block_max(em, scratch);
    if (tid == 0) s_e2max = em;
    __syncthreads();
    const double scale = fmax(fabs(s_lo), fabs(s_hi));
    const double pad = 2.220446049250313e-16 * n * scale + 1e-300;
    double lo = s_lo - pad, hi = s_hi + pad;
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, s_e2max) * 4.0;
    const int idx = n - 1 - (int)blockIdx.x;             // ascending index of the wanted eigenvalue
    for (int pass = 0; pass < passes; pass++) {
        const double step = (hi - lo) * (1.0 / (BIS + 1));
        const double x = fma(step, (double)(tid + 1), lo);
        int cnt = 0;
        double q = d[0] - x;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0;
        for (int i = 1; i < n; i++) {
            q = (d[i] - x) - e2[i - 1] * rcp_fast(q);
            if (fabs(q) < pivmin) q = -pivmin;
            cnt += q < 0.0;
        }
        const int L = __syncthreads_count(cnt <= idx);   // points still at or below the eigenvalue
        const double nlo = L > 0 ? fma(step, (double)L, lo) : lo;
        const double nhi = L < BIS ? fma(step, (double)(L + 1), lo) : hi;
        lo = nlo;
        hi = nhi;
    }
    if (tid == 0) {
        lam[blockIdx.x] = 0.5 * (lo + hi);
        if (blockIdx.x == 0) bounds[0] = scale;
    }
}

// ---------------------------------------------------------------------------------------------
// 3. One step of inverse iteration for shift t: x = (T - lam_t I)^-1 b, normalised.  Pivoted
// LU of the shifted tridiagonal (row interchanges as LAPACK's gttrf), b = row t of Bt or a
// fixed pseudo-random vector.  The recurrences are sequential: lane 0 runs them out of shared
// memory, the warp loads, normalises and stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double hash_unit(unsigned a, unsigned b) {
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return ((double)(h >> 8) + 0.5) * (2.0 / 16777216.0) - 1.0;   // (-1, 1), never 0
}

__global__ void __launch_bounds__(32)
invit_kernel(const double* __restrict__ d_g, const double* __restrict__ e_g, int n, const double* __restrict__ lam,
             const double* __restrict__ bounds, const double* __restrict__ Bt, double* __restrict__ Xt) {
    extern __shared__ double sm[];
    double* dd = sm;               // d - mu, later the reciprocal pivots
    double* ee = dd + n;           // off-diagonal
    double* L = ee + n;            // multipliers
    double* U1 = L + n;            // first super-diagonal of U
    double* U2 = U1 + n;           // second super-diagonal of U
    double* x = U2 + n;            // right-hand side / solution
    unsigned char* piv = reinterpret_cast<unsigned char*>(x + n);
    const int lane = threadIdx.x, t = blockIdx.x;
    const double mu = lam[t];
    const double tiny = fmax(bounds[0] * 2.220446049250313e-16, 1e-290);
    for (int i = lane; i < n; i += 32) {
        dd[i] = d_g[i] - mu;
        ee[i] = i < n - 1 ? e_g[i] : 0.0;
        x[i] = Bt ? Bt[(size_t)t * n + i] : hash_unit((unsigned)t, (unsigned)i);
    }
    __syncwarp();
    if (lane == 0) {
        double dcur = dd[0], ucur = ee[0];
        for (int i = 0; i < n - 1; i++) {
            const double sub = ee[i], dnext = dd[i + 1], unext = ee[i + 1];   // ee[n-1] = 0
            if (fabs(dcur) >= fabs(sub) || fabs(sub) < tiny) {
                if (fabs(dcur) < tiny) dcur = dcur < 0.0 ? -tiny : tiny;
                const double inv = rcp_fast(dcur), f = sub * inv;
                L[i] = f; piv[i] = 0; dd[i] = inv; U1[i] = ucur; U2[i] = 0.0;
                dcur = fma(-f, ucur, dnext);
                ucur = unext;
            } else {
                const double inv = rcp_fast(sub), f = dcur * inv;
                L[i] = f; piv[i] = 1; dd[i] = inv; U1[i] = dnext; U2[i] = unext;
                dcur = fma(-f, dnext, ucur);
                ucur = -f * unext;
            }
        }
        if (fabs(dcur) < tiny) dcur = dcur < 0.0 ? -tiny : tiny;
        dd[n - 1] = rcp_fast(dcur);
        // forward substitution with the interchanges
        double cur = x[0];
        for (int i = 0; i < n - 1; i++) {
            const double nxt = x[i + 1];
            if (!piv[i]) { x[i] = cur; cur = fma(-L[i], cur, nxt); }
            else { x[i] = nxt; cur = fma(-L[i], nxt, cur); }
        }
        // back substitution
        double x1 = cur * dd[n - 1], x2 = 0.0;
        x[n - 1] = x1;
        for (int i = n - 2; i >= 0; i--) {
            double v = x[i];
            v = fma(-U1[i], x1, v);
            v = fma(-U2[i], x2, v);
            v *= dd[i];
            x[i] = v;
            x2 = x1;
            x1 = v;
        }
    }
    __syncwarp();
    double big = 0.0;
    for (int i = lane; i < n; i += 32) big = fmax(big, fabs(x[i]));
    big = warp_max(big);
    const double sc = big > 0.0 ? 1.0 / big : 1.0;       // scale first: |x| can be ~1e16 / |T|
    double ss = 0.0;
    for (int i = lane; i < n; i += 32) { const double v = x[i] * sc; ss = fma(v, v, ss); }
    ss = warp_sum(ss);
    const double nrm = ss > 0.0 ? sc / sqrt(ss) : 0.0;
    for (int i = lane; i < n; i += 32) Xt[(size_t)t * n + i] = x[i] * nrm;
}

// out[a][b] = sum_i A[a][i] B[b][i]   (A: ma x n, B: mb x n, both row-major); one CTA per row a
__global__ void __launch_bounds__(256)
rows_dot_kernel(const double* __restrict__ A, const double* __restrict__ B, int mb, int n, double* __restrict__ out) {
    extern __shared__ double sm[];
    const int a = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n; i += 256) sm[i] = A[(size_t)a * n + i];
    __syncthreads();
    for (int b = warp; b < mb; b += 8) {
        const double* row = B + (size_t)b * n;
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) acc = fma(sm[i], row[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[(size_t)a * mb + b] = acc;
    }
}

// out[c][i] = sum_r coef(c, r) X[r][i],  coef(c, r) = Cf[c * crs + r * ccs];  m <= 128
// grid (ceil(n / 128), ceil(m / 8)), 128 threads: a thread owns one i and 8 consecutive c
__global__ void __launch_bounds__(128)
combine_rows_kernel(const double* __restrict__ Cf, int64_t crs, int64_t ccs, const double* __restrict__ X, int m, int n,
                    double* __restrict__ out) {
    __shared__ double cf[8][128];
    const int c0 = blockIdx.y * 8, i = blockIdx.x * 128 + threadIdx.x;
    for (int e = threadIdx.x; e < 8 * m; e += 128) {
        const int cc = e / m, r = e - cc * m;
        cf[cc][r] = c0 + cc < m ? Cf[(size_t)(c0 + cc) * crs + (size_t)r * ccs] : 0.0;
    }
    __syncthreads();
    if (i >= n) return;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < m; r++) {
        const double xv = X[(size_t)r * n + i];
#pragma unroll
        for (int cc = 0; cc < 8; cc++) acc[cc] = fma(cf[cc][r], xv, acc[cc]);
    }
#pragma unroll
    for (int cc = 0; cc < 8; cc++)
        if (c0 + cc < m) out[(size_t)(c0 + cc) * n + i] = acc[cc];
}

// S = L L^T (m <= 128, one CTA), Linv = L^-1 (lower triangular, row-major).  A pivot that drops
// below m eps of its original diagonal means the block lost rank numerically: counted in
// info[0] (the caller falls back to the full solver) and regularised so the kernel terminates.
__global__ void __launch_bounds__(256)
chol_inverse_kernel(const double* __restrict__ S, int m, double* __restrict__ Linv, int* __restrict__ info) {
    extern __shared__ double sm[];
    double* Lm = sm;                 // m x (m + 1)
    double* Y = sm + (size_t)m * (m + 1);   // m x (m + 1)
    __shared__ int s_bad;
    const int tid = threadIdx.x, ld = m + 1;
    if (tid == 0) s_bad = 0;
    for (int e = tid; e < m * m; e += 256) {
        const int r = e / m, c = e - r * m;
        Lm[r * ld + c] = 0.5 * (S[(size_t)r * m + c] + S[(size_t)c * m + r]);
    }
    __syncthreads();
    for (int j = 0; j < m; j++) {
        double piv = Lm[j * ld + j];
        const double floor_j = 2.220446049250313e-16 * m * fabs(S[(size_t)j * m + j]);
        if (!(piv > floor_j)) {
            if (tid == 0) s_bad++;
            piv = floor_j > 0.0 ? floor_j : 1e-300;
        }
        const double root = sqrt(piv), inv = 1.0 / root;
        __syncthreads();
        for (int r = j + tid; r < m; r += 256) Lm[r * ld + j] = r == j ? root : Lm[r * ld + j] * inv;
        __syncthreads();
        const int rem = m - j - 1;
        for (int e = tid; e < rem * rem; e += 256) {
            const int r = j + 1 + e / rem, c = j + 1 + e % rem;
            if (c <= r) Lm[r * ld + c] = fma(-Lm[r * ld + j], Lm[c * ld + j], Lm[r * ld + c]);
        }
        __syncthreads();
    }
    // column c of Y = L^-1 e_c by forward substitution, one thread per column
    for (int c = tid; c < m; c += 256) {
        for (int r = 0; r < m; r++) {
            double v = r == c ? 1.0 : 0.0;
            if (r < c) { Y[r * ld + c] = 0.0; continue; }
            for (int q = c; q < r; q++) v = fma(-Lm[r * ld + q], Y[q * ld + c], v);
            Y[r * ld + c] = v / Lm[r * ld + r];
        }
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += 256) {
        const int r = e / m, c = e - r * m;
        Linv[e] = Y[r * ld + c];
    }
    if (tid == 0) atomicAdd(info, s_bad);
}

// Y[t][i] = (T X[t])_i
__global__ void __launch_bounds__(256)
tridiag_apply_kernel(const double* __restrict__ d, const double* __restrict__ e, const double* __restrict__ X, int m, int n,
                     double* __restrict__ Y) {
    const int64_t total = (int64_t)m * n;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int i = (int)(idx % n);
        double v = d[i] * X[idx];
        if (i > 0) v = fma(e[i - 1], X[idx - 1], v);
        if (i < n - 1) v = fma(e[i], X[idx + 1], v);
        Y[idx] = v;
    }
}

__global__ void __launch_bounds__(256) symmetrise_kernel(double* H, int m) {
    for (int e = blockIdx.x * 256 + threadIdx.x; e < m * m; e += gridDim.x * 256) {
        const int r = e / m, c = e - r * m;
        if (c > r) {
            const double v = 0.5 * (H[e] + H[(size_t)c * m + r]);
            H[e] = v;
            H[(size_t)c * m + r] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 5. Back-transformation: one warp per Ritz vector, z <- H_0 H_1 ... H_{n-3} z with the
// reflector rows streamed from L2 one step ahead of their use.  U[i * ldu + t] = z_i.
// ---------------------------------------------------------------------------------------------
template <int NE>
__global__ void __launch_bounds__(128)
backtransform_kernel(const double* __restrict__ V, const double* __restrict__ tau_g, int n, const double* __restrict__ Zt,
                     int m, double* __restrict__ U, int64_t ldu) {
    const int lane = threadIdx.x & 31, t = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= m) return;
    double z[NE], va[NE], vb[NE];
#pragma unroll
    for (int q = 0; q < NE; q++) {
        const int k = lane + 32 * q;
        z[q] = k < n ? Zt[(size_t)t * n + k] : 0.0;
    }
    auto load = [&](double* dst, int j) {
        const double* v = V + (size_t)j * n;
#pragma unroll
        for (int q = 0; q < NE; q++) {
            const int k = lane + 32 * q;
            dst[q] = (k > j && k < n) ? __ldg(v + k) : 0.0;
        }
    };
    auto apply = [&](const double* v, double tau) {
        double dot = 0.0;
#pragma unroll
        for (int q = 0; q < NE; q++) dot = fma(v[q], z[q], dot);
        dot = warp_sum(dot) * tau;
#pragma unroll
        for (int q = 0; q < NE; q++) z[q] = fma(-dot, v[q], z[q]);
    };
    int j = n - 3;
    if (j >= 0) load(va, j);
    while (j >= 0) {
        if (j >= 1) load(vb, j - 1);
        apply(va, tau_g[j]);
        if (j >= 1) {
            if (j >= 2) load(va, j - 2);
            apply(vb, tau_g[j - 1]);
        }
        j -= 2;
    }
#pragma unroll
    for (int q = 0; q < NE; q++) {
        const int k = lane + 32 * q;
        if (k < n) U[(size_t)k * ldu + t] = z[q];
    }
}

// out[0..k-1] = Ritz values (descending), out[k] = trace(G), out[k+1] = rank-loss count of the
// Cholesky-QR steps (0 when healthy)
__global__ void __launch_bounds__(256)
finish_kernel(const double* __restrict__ G, int n, const double* __restrict__ hev, int k, const int* __restrict__ info,
              double* __restrict__ out) {
    __shared__ double scratch[32];
    double tr = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) tr += G[(size_t)i * n + i];
    tr = block_sum(tr, scratch);
    for (int i = threadIdx.x; i < k; i += 256) out[i] = hev[i];
    if (threadIdx.x == 0) { out[k] = tr; out[k + 1] = (double)info[0]; }
}

static int orthonormalise(ndmps_ctx* ctx, const double* Xin, int m, int n, double* S, double* Linv, int* info, double* Xout) {
    rows_dot_kernel<<<m, 256, (size_t)n * sizeof(double), ctx->stream>>>(Xin, Xin, m, n, S);
    NDMPS_LAUNCH_CHECK(ctx);
    chol_inverse_kernel<<<1, 256, (size_t)2 * m * (m + 1) * sizeof(double), ctx->stream>>>(S, m, Linv, info);
    NDMPS_LAUNCH_CHECK(ctx);
    combine_rows_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)((m + 7) / 8)), 128, 0, ctx->stream>>>(Linv, m, 1, Xin, m, n, Xout);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}

}  // namespace topk

// Leading k eigenpairs of the symmetric positive semi-definite G (n x n, float64, device; not
// modified).  out_dev: k + 2 doubles (see finish_kernel); U: eigenvector t in column t, leading
// dimension ldu >= k.  *done = false (nothing computed) when the shape is outside what this
// path supports.
int eigh_topk(ndmps_ctx* ctx, const double* G, int64_t n64, int64_t k64, double* out_dev, double* U, int64_t ldu, bool* done) {
    using namespace topk;
    *done = false;
    if (n64 < 96 || n64 > TDT * KM || k64 < 1 || k64 > 128 || 2 * k64 > n64) return NDMPS_OK;
    const int n = (int)n64, m = (int)k64;
    int C = (n + 7) / 8;                                 // 8 rows per CTA: one warp per row
    const int cmax = ctx->sm_count < 128 ? ctx->sm_count : 128;
    if (ctx->opt_topk_rows > 0) C = (int)((n + ctx->opt_topk_rows - 1) / ctx->opt_topk_rows);
    if (C > cmax) C = cmax;
    if (C < 1) C = 1;
    int nr = (n + C - 1) / C;
    const size_t smem_td = ((size_t)nr * n + 3 * (size_t)n) * sizeof(double);
    const size_t smem_iv = (size_t)6 * n * sizeof(double) + (size_t)n + 16;
    const size_t smem_ch = (size_t)2 * m * (m + 1) * sizeof(double);
    if (smem_td + 2048 > ctx->smem_optin || smem_ch + 2048 > ctx->smem_optin) return NDMPS_OK;
    {
        static size_t td_set = 0, iv_set = 0, ch_set = 0;
        if (smem_td > td_set) {
            NDMPS_CUDA_TRY(cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_td));
            td_set = smem_td;
        }
        if (smem_iv > iv_set) {
            NDMPS_CUDA_TRY(cudaFuncSetAttribute(invit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_iv));
            iv_set = smem_iv;
        }
        if (smem_ch > ch_set) {
            NDMPS_CUDA_TRY(cudaFuncSetAttribute(chol_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ch));
            ch_set = smem_ch;
        }
    }
    double *V, *tau, *d, *e, *pbuf, *rowbuf, *lam, *bounds, *Xa, *Xb, *S, *Linv, *TQ, *H, *hev, *W;
    unsigned* ctrl;
    int* info;
    NDMPS_TRY(ctx->ws.get<double>((size_t)n * n, &V));
    NDMPS_TRY(ctx->ws.get<double>((size_t)n, &tau));
    NDMPS_TRY(ctx->ws.get<double>((size_t)n, &d));
    NDMPS_TRY(ctx->ws.get<double>((size_t)n, &e));
    NDMPS_TRY(ctx->ws.get<double>((size_t)2 * n, &pbuf));
    NDMPS_TRY(ctx->ws.get<double>((size_t)2 * n, &rowbuf));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m, &lam));
    NDMPS_TRY(ctx->ws.get<double>(4, &bounds));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m * n, &Xa));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m * n, &Xb));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m * m, &S));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m * m, &Linv));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m * n, &TQ));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m * m, &H));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m, &hev));
    NDMPS_TRY(ctx->ws.get<double>((size_t)m * m, &W));
    NDMPS_TRY(ctx->ws.get<unsigned>(8, &ctrl));
    info = reinterpret_cast<int*>(ctrl + 4);
    NDMPS_CUDA_TRY(cudaMemsetAsync(ctrl, 0, 8 * sizeof(unsigned), ctx->stream));

    // 1. tridiagonalise
    {
        int n_arg = n, nr_arg = nr;
        void* args[] = {(void*)&G, &n_arg, &nr_arg, &V, &tau, &d, &e, &pbuf, &rowbuf, &ctrl};
        NDMPS_CUDA_TRY(cudaLaunchCooperativeKernel((void*)tridiag_kernel, dim3(C), dim3(TDT), args, smem_td, ctx->stream));
        ctx->launches++;
    }
    // 2. shifts
    const int passes = ctx->opt_topk_passes > 0 ? (int)ctx->opt_topk_passes : 8;
    bisect_kernel<<<m, BIS, (size_t)2 * n * sizeof(double), ctx->stream>>>(d, e, n, passes, lam, bounds);
    NDMPS_LAUNCH_CHECK(ctx);
    // 3. inverse iteration, re-orthonormalised between steps; the last basis is orthonormalised twice
    const int iters = ctx->opt_topk_iters > 0 ? (int)ctx->opt_topk_iters : 3;
    const double* rhs = nullptr;
    for (int it = 0; it < iters; it++) {
        invit_kernel<<<m, 32, smem_iv, ctx->stream>>>(d, e, n, lam, bounds, rhs, Xa);
        NDMPS_LAUNCH_CHECK(ctx);
        NDMPS_TRY(orthonormalise(ctx, Xa, m, n, S, Linv, info, Xb));
        rhs = Xb;
    }
    NDMPS_TRY(orthonormalise(ctx, Xb, m, n, S, Linv, info, Xa));       // Q = Xa
    // 4. Rayleigh-Ritz on span(Q)
    tridiag_apply_kernel<<<(unsigned)(((int64_t)m * n + 255) / 256), 256, 0, ctx->stream>>>(d, e, Xa, m, n, TQ);
    NDMPS_LAUNCH_CHECK(ctx);
    rows_dot_kernel<<<m, 256, (size_t)n * sizeof(double), ctx->stream>>>(Xa, TQ, m, n, H);
    NDMPS_LAUNCH_CHECK(ctx);
    symmetrise_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, ctx->stream>>>(H, m);
    NDMPS_LAUNCH_CHECK(ctx);
    NDMPS_TRY(eigh(ctx, H, m, hev, W, 0.0));
    // Zt[c] = sum_r W[r][c] Q[r]
    combine_rows_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)((m + 7) / 8)), 128, 0, ctx->stream>>>(W, 1, m, Xa, m, n, Xb);
    NDMPS_LAUNCH_CHECK(ctx);
    // 5. back-transform
    if (n <= 512) backtransform_kernel<16><<<(unsigned)((m + 3) / 4), 128, 0, ctx->stream>>>(V, tau, n, Xb, m, U, ldu);
    else backtransform_kernel<32><<<(unsigned)((m + 3) / 4), 128, 0, ctx->stream>>>(V, tau, n, Xb, m, U, ldu);
    NDMPS_LAUNCH_CHECK(ctx);
    finish_kernel<<<1, 256, 0, ctx->stream>>>(G, n, hev, m, info, out_dev);
    NDMPS_LAUNCH_CHECK(ctx);
    if (ctx->opt_verbose) fprintf(stderr, "[ndmps] eigh_topk n = %d, k = %d: %d CTAs x %d rows\n", n, m, C, nr);
    *done = true;
    return NDMPS_OK;
}

}  // namespace ndmps

using namespace ndmps;

extern "C" {

int ndmps_eigh_topk(ndmps_ctx_t* ctx, const double* a_dev, int64_t n, int64_t k, double* evals_dev, double* evecs_dev) {
    NDMPS_REQUIRE(ctx && a_dev && evals_dev && evecs_dev, "ndmps_eigh_topk: NULL argument");
    NDMPS_REQUIRE(n >= 1 && k >= 1 && k <= n, "ndmps_eigh_topk: need 1 <= k <= n");
    NDMPS_TRY(ctx->ws.reset(ctx->stream));
    bool done = false;
    NDMPS_TRY(eigh_topk(ctx, a_dev, n, k, evals_dev, evecs_dev, k, &done));
    if (!done) {
        set_error("ndmps_eigh_topk: shape n = %lld, k = %lld is outside the leading-eigenpair path (96 <= n <= 1024, 2k <= n, k <= 128)",
                  (long long)n, (long long)k);
        return NDMPS_ERR_INVALID;
    }
    return NDMPS_OK;
}

}  // extern "C"
