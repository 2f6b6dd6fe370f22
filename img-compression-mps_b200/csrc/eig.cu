// K3: symmetric eigen-decomposition of bond-sized Gram matrices, float64,
// by parallel one-sided (Hestenes) Jacobi.
//
// Replaces the LAPACK gesdd call quimb makes under from_dense / tensor_compress_bond
// (core/ndmps.py:74,106) for the small matrix of each sweep step: with
// G = M M^T symmetric positive semi-definite, orthogonalising the columns of G by
// plane rotations gives G V = W with W's columns mutually orthogonal, so
// lambda_j = |w_j| and u_j = w_j / |w_j|.  Everything is float64 (B200's FP64 pipe),
// because the sweep's rank decisions sit at lambda/lambda_max ~ 1e-10.
//
// Parallel scheme: the n columns are split into nb blocks of b columns.  A sweep
// is a round-robin tournament over blocks (nb-1 rounds); in a round every CTA owns
// one block pair in shared memory and one warp owns one column pair at a time.
// Round 0 of a sweep runs the full tournament inside the 2b columns (this covers
// the within-block pairs); later rounds only rotate cross pairs (i in P, j in Q).
// Matrices that fit one CTA's shared memory are swept to convergence in a single
// launch.
#include "common.cuh"

namespace ndmps {

// rotation threshold |x.y| <= tol |x||y| with tol = sqrt(n) * 2 eps (LAPACK dgesvj uses sqrt(n) eps)
static inline double jacobi_tol(int n) { return sqrt((double)n) * 4.4408920985006262e-16; }

// rotate columns x, y (length n, shared memory) so that they become orthogonal.
// NR > 0: n <= 32*NR, columns staged in registers.  NR == 0: generic n, two passes.
template <int NR>
__device__ __forceinline__ bool rotate_pair(double* __restrict__ x, double* __restrict__ y, int n, int lane,
                                            double tol, double floor2) {
    double xr[NR > 0 ? NR : 1], yr[NR > 0 ? NR : 1];
    double app = 0.0, aqq = 0.0, apq = 0.0;
    if (NR > 0) {
#pragma unroll
        for (int t = 0; t < NR; t++) {
            int i = lane + 32 * t;
            xr[t] = i < n ? x[i] : 0.0;
            yr[t] = i < n ? y[i] : 0.0;
            app = fma(xr[t], xr[t], app);
            aqq = fma(yr[t], yr[t], aqq);
            apq = fma(xr[t], yr[t], apq);
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            double a = x[i], b = y[i];
            app = fma(a, a, app);
            aqq = fma(b, b, aqq);
            apq = fma(a, b, apq);
        }
    }
    app = warp_sum(app);
    aqq = warp_sum(aqq);
    apq = warp_sum(apq);
    // columns whose norm is below n*eps*|G| are numerically null: rotating them only churns
    // round-off and would keep the sweep from ever reporting convergence
    if (app <= floor2 || aqq <= floor2) return false;
    if (apq == 0.0 || fabs(apq) <= tol * sqrt(app) * sqrt(aqq)) return false;
    double zeta = (aqq - app) / (2.0 * apq);
    double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    double c = 1.0 / sqrt(1.0 + t * t);
    double s = c * t;
    if (NR > 0) {
#pragma unroll
        for (int tt = 0; tt < NR; tt++) {
            int i = lane + 32 * tt;
            if (i < n) {
                x[i] = c * xr[tt] - s * yr[tt];
                y[i] = s * xr[tt] + c * yr[tt];
            }
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            double a = x[i], b = y[i];
            x[i] = c * a - s * b;
            y[i] = s * a + c * b;
        }
    }
    return true;
}

// slot pair of a round-robin tournament over P (even) players, match `w` of round `lr`
__device__ __forceinline__ void tournament_pair(int P, int lr, int w, int& s1, int& s2) {
    int m = P - 1;
    if (w == 0) {
        s1 = m;
        s2 = lr;
    } else {
        s1 = (lr + w) % m;
        s2 = (lr - w + m) % m;
    }
}

// One round of the block tournament.  grid = nb/2 CTAs, block = 32*b threads,
// dynamic smem = 2*b*n doubles.  A is n x n with "column" j stored at A + j*n
// (the input is symmetric, so row-major == column-major).
template <int NR>
__global__ void __launch_bounds__(512)
jacobi_round_kernel(double* __restrict__ A, int n, int b, int nb, int round, int* __restrict__ rotated, double tol,
                    const double* __restrict__ floor2_ptr) {
    extern __shared__ double S[];
    const double floor2 = *floor2_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int p, q;
    tournament_pair(nb, round, blockIdx.x, p, q);
    const int colp0 = p * b, colq0 = q * b;
    int cntp = n - colp0; cntp = cntp < 0 ? 0 : (cntp > b ? b : cntp);
    int cntq = n - colq0; cntq = cntq < 0 ? 0 : (cntq > b ? b : cntq);
    // load the two column blocks
    for (int lc = warp; lc < 2 * b; lc += b) {
        bool isq = lc >= b;
        int l = isq ? lc - b : lc;
        if (l < (isq ? cntq : cntp)) {
            const double* src = A + (size_t)((isq ? colq0 : colp0) + l) * n;
            double* dst = S + (size_t)lc * n;
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
        }
    }
    __syncthreads();
    bool any = false;
    if (round == 0) {
        const int P = 2 * b;
        for (int lr = 0; lr < P - 1; lr++) {
            int s1, s2;
            tournament_pair(P, lr, warp, s1, s2);
            bool v1 = s1 < b ? s1 < cntp : (s1 - b) < cntq;
            bool v2 = s2 < b ? s2 < cntp : (s2 - b) < cntq;
            if (v1 && v2) any |= rotate_pair<NR>(S + (size_t)s1 * n, S + (size_t)s2 * n, n, lane, tol, floor2);
            __syncthreads();
        }
    } else {
        for (int k = 0; k < b; k++) {
            int j = warp + k; j = j >= b ? j - b : j;
            if (warp < cntp && j < cntq) any |= rotate_pair<NR>(S + (size_t)warp * n, S + (size_t)(b + j) * n, n, lane, tol, floor2);
            __syncthreads();
        }
    }
    for (int lc = warp; lc < 2 * b; lc += b) {
        bool isq = lc >= b;
        int l = isq ? lc - b : lc;
        if (l < (isq ? cntq : cntp)) {
            double* dst = A + (size_t)((isq ? colq0 : colp0) + l) * n;
            const double* src = S + (size_t)lc * n;
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
        }
    }
    if (any && lane == 0) *rotated = 1;
}

// Whole matrix in one CTA: sweep until no rotation happens.  block = 32*W threads,
// dynamic smem = n*n doubles.  sweeps_out[0] = sweeps used (negative: not converged).
template <int NR>
__global__ void __launch_bounds__(1024)
jacobi_single_kernel(double* __restrict__ A, int n, int max_sweeps, int* __restrict__ sweeps_out, double tol,
                     const double* __restrict__ floor2_ptr) {
    extern __shared__ double S[];
    const double floor2 = *floor2_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) S[i] = A[i];
    __syncthreads();
    const int P = (n + 1) & ~1;          // pad to an even number of players; player n (if any) is a bye
    const int matches = P / 2;
    int sweep = 0;
    int done = n < 2 ? 1 : 0;
    while (!done && sweep < max_sweeps) {
        bool any = false;
        for (int lr = 0; lr < P - 1; lr++) {
            for (int w = warp; w < matches; w += W) {
                int s1, s2;
                tournament_pair(P, lr, w, s1, s2);
                if (s1 < n && s2 < n) any |= rotate_pair<NR>(S + (size_t)s1 * n, S + (size_t)s2 * n, n, lane, tol, floor2);
            }
            __syncthreads();
        }
        sweep++;
        done = !__syncthreads_or(any ? 1 : 0);
    }
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) A[i] = S[i];
    if (threadIdx.x == 0) sweeps_out[0] = done ? sweep : -sweep;
}

// lambda_j = |column j|, one warp per column
__global__ void __launch_bounds__(256) column_norms_kernel(const double* __restrict__ A, int n, double* __restrict__ norms) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double* col = A + (size_t)warp * n;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s = fma(col[i], col[i], s);
    s = warp_sum(s);
    if (lane == 0) norms[warp] = sqrt(s);
}

// floor2 = (n * eps * max_j |column j|)^2 : squared norm below which a column counts as null
__global__ void __launch_bounds__(256) null_floor_kernel(const double* __restrict__ norms, int n, double* __restrict__ floor2) {
    __shared__ double scratch[32];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, norms[i]);
    m = block_max(m, scratch);
    if (threadIdx.x == 0) {
        double f = (double)n * 2.220446049250313e-16 * m;
        floor2[0] = f * f;
    }
}

// descending rank of every column by counting (ties broken by index), then write
// evals[rank] and the normalised column into evecs[:, rank] (row-major n x n)
__global__ void __launch_bounds__(256)
sort_extract_kernel(const double* __restrict__ A, const double* __restrict__ norms, int n, double* __restrict__ evals,
                    double* __restrict__ evecs) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    double mine = norms[warp];
    int cnt = 0;
    for (int k = lane; k < n; k += 32) {
        double o = norms[k];
        cnt += (o > mine || (o == mine && k < warp)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    int rank = cnt;
    if (lane == 0) evals[rank] = mine;
    double inv = mine > 0.0 ? 1.0 / mine : 0.0;
    const double* col = A + (size_t)warp * n;
    for (int i = lane; i < n; i += 32) evecs[(size_t)i * n + rank] = col[i] * inv;
}

template <int NR>
static int launch_round(ndmps_ctx* ctx, double* A, int n, int b, int nb, int round, int* flag, size_t smem,
                        const double* floor2) {
    static bool attr_set = false;
    if (!attr_set) {
        NDMPS_CUDA_TRY(cudaFuncSetAttribute(jacobi_round_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)ctx->smem_optin));
        attr_set = true;
    }
    jacobi_round_kernel<NR><<<nb / 2, 32 * b, smem, ctx->stream>>>(A, n, b, nb, round, flag, jacobi_tol(n), floor2);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}

template <int NR>
static int launch_single(ndmps_ctx* ctx, double* A, int n, int warps, int max_sweeps, int* sweeps_dev, size_t smem,
                         const double* floor2) {
    static bool attr_set = false;
    if (!attr_set) {
        NDMPS_CUDA_TRY(cudaFuncSetAttribute(jacobi_single_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)ctx->smem_optin));
        attr_set = true;
    }
    jacobi_single_kernel<NR><<<1, 32 * warps, smem, ctx->stream>>>(A, n, max_sweeps, sweeps_dev, jacobi_tol(n), floor2);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}

int eigh(ndmps_ctx* ctx, double* a_dev, int64_t n64, double* evals_dev, double* evecs_dev) {
    NDMPS_REQUIRE(n64 >= 1 && n64 <= 16384, "eigh: n = %lld outside 1..16384", (long long)n64);
    const int n = (int)n64;
    const int max_sweeps = (int)ctx->opt_jacobi_max_sweeps;
    const size_t smem_cap = ctx->smem_optin > 4096 ? ctx->smem_optin - 1024 : 0;
    NDMPS_TRY(ensure_pinned(ctx, 64));
    int* flags = nullptr;   // device ints: [0] single-kernel sweeps, [1..] per-sweep rotation flags
    NDMPS_TRY(ctx->ws.get<int>((size_t)max_sweeps + 2, &flags));
    NDMPS_CUDA_TRY(cudaMemsetAsync(flags, 0, ((size_t)max_sweeps + 2) * sizeof(int), ctx->stream));
    int* host_flag = reinterpret_cast<int*>(ctx->pinned);
    int sweeps_used = 0;
    double* norms = nullptr;
    double* floor2 = nullptr;
    NDMPS_TRY(ctx->ws.get<double>((size_t)n, &norms));
    NDMPS_TRY(ctx->ws.get<double>(1, &floor2));
    const int ngrid = (n * 32 + 255) / 256;
    column_norms_kernel<<<ngrid, 256, 0, ctx->stream>>>(a_dev, n, norms);
    NDMPS_LAUNCH_CHECK(ctx);
    null_floor_kernel<<<1, 256, 0, ctx->stream>>>(norms, n, floor2);
    NDMPS_LAUNCH_CHECK(ctx);
    const int nr = n <= 64 ? 2 : n <= 128 ? 4 : n <= 256 ? 8 : n <= 512 ? 16 : 0;

    const size_t single_bytes = (size_t)n * n * sizeof(double);
    if (n >= 2 && single_bytes <= smem_cap && n <= 128) {
        int matches = (n + 1) / 2;
        int warps = matches < 32 ? matches : 32;
        switch (nr) {
            case 2: NDMPS_TRY(launch_single<2>(ctx, a_dev, n, warps, max_sweeps, flags, single_bytes, floor2)); break;
            default: NDMPS_TRY(launch_single<4>(ctx, a_dev, n, warps, max_sweeps, flags, single_bytes, floor2)); break;
        }
        NDMPS_CUDA_TRY(cudaMemcpyAsync(host_flag, flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        NDMPS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (host_flag[0] <= 0) {
            set_error("eigh: Jacobi did not converge in %d sweeps (n = %d)", max_sweeps, n);
            return NDMPS_ERR_NOCONV;
        }
        sweeps_used = host_flag[0];
    } else if (n >= 2) {
        // block size: as many columns as shared memory allows, at most 32 warps, tunable
        int b = (int)(smem_cap / (16 * (size_t)n));
        if (b > 16) b = 16;
        if (ctx->opt_jacobi_block > 0 && ctx->opt_jacobi_block < b) b = (int)ctx->opt_jacobi_block;
        NDMPS_REQUIRE(b >= 1, "eigh: n = %d does not fit a column pair in shared memory", n);
        int nb = (n + b - 1) / b;
        if (nb & 1) nb++;
        if (nb < 2) nb = 2;
        size_t smem = (size_t)2 * b * n * sizeof(double);
        bool converged = false;
        for (int s = 0; s < max_sweeps && !converged; s++) {
            int* flag = flags + 1 + s;
            for (int round = 0; round < nb - 1; round++) {
                switch (nr) {
                    case 4: NDMPS_TRY(launch_round<4>(ctx, a_dev, n, b, nb, round, flag, smem, floor2)); break;
                    case 8: NDMPS_TRY(launch_round<8>(ctx, a_dev, n, b, nb, round, flag, smem, floor2)); break;
                    case 16: NDMPS_TRY(launch_round<16>(ctx, a_dev, n, b, nb, round, flag, smem, floor2)); break;
                    default: NDMPS_TRY(launch_round<0>(ctx, a_dev, n, b, nb, round, flag, smem, floor2)); break;
                }
            }
            sweeps_used = s + 1;
            if (s >= 3) {
                NDMPS_CUDA_TRY(cudaMemcpyAsync(host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                NDMPS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                converged = host_flag[0] == 0;
            }
        }
        if (!converged) {
            set_error("eigh: Jacobi did not converge in %d sweeps (n = %d, b = %d)", max_sweeps, n, b);
            return NDMPS_ERR_NOCONV;
        }
    }
    ctx->last_eig_sweeps = sweeps_used;
    int grid = ngrid;
    column_norms_kernel<<<grid, 256, 0, ctx->stream>>>(a_dev, n, norms);
    NDMPS_LAUNCH_CHECK(ctx);
    sort_extract_kernel<<<grid, 256, 0, ctx->stream>>>(a_dev, norms, n, evals_dev, evecs_dev);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}

}  // namespace ndmps

using namespace ndmps;

extern "C" {

int ndmps_eigh(ndmps_ctx_t* ctx, double* a_dev, int64_t n, double* evals_dev, double* evecs_dev, int* sweeps_out_host) {
    NDMPS_REQUIRE(ctx && a_dev && evals_dev && evecs_dev, "ndmps_eigh: NULL argument");
    NDMPS_TRY(ctx->ws.reset(ctx->stream));
    NDMPS_TRY(eigh(ctx, a_dev, n, evals_dev, evecs_dev));
    if (sweeps_out_host) *sweeps_out_host = ctx->last_eig_sweeps;
    return NDMPS_OK;
}

}  // extern "C"
