// K1: N-D volume <-> interleaved MPS site order.
//
// The reference builds an int64 map (utils/core.py:6-35,129-168) and scatters /
// gathers through it (core/ndmps.py:66-71,144-148).  The map is a pure
// permutation: axis a of the volume splits into per-level digits f[l][a]
// (most significant first), and site l is the C-order fusion of the level-l
// digits of all axes.  So   dst_digits(l, a)  <->  src_digits(a, l).
//
// Here the plan stores, for each direction, the destination's digit extents (in
// C order) and the source stride of every digit; kernels walk destination
// offsets (coalesced writes) and gather the source.
#include <type_traits>

#include "common.cuh"

namespace ndmps {

static void push_digit(DigitList& d, uint32_t extent, int64_t stride) {
    if (extent == 1) return;
    if (d.n > 0 && d.stride[d.n - 1] == stride * (int64_t)extent &&
        (uint64_t)d.extent[d.n - 1] * extent < (1ull << 31)) {
        // previous (outer) digit is contiguous with this one in the source: fuse
        d.extent[d.n - 1] *= extent;
        d.stride[d.n - 1] = stride;
    } else {
        d.extent[d.n] = extent;
        d.stride[d.n] = stride;
        d.n++;
    }
}

static void finish(DigitList& d) {
    if (d.n == 0) { d.extent[0] = 1; d.stride[0] = 0; d.n = 1; }
    for (int i = 0; i < d.n; i++) {
        uint32_t e = d.extent[i];
        d.shift[i] = (e & (e - 1)) == 0 ? __builtin_ctz(e) : -1;
    }
}

// source offset of destination element `o`: peel destination digits from the
// innermost outwards (shared by the kernels and the host-side plan check)
template <class idx_t>
__host__ __device__ __forceinline__ int64_t digit_offset(const DigitList& dl, idx_t rem) {
    int64_t off = 0;
#pragma unroll 1
    for (int j = dl.n - 1; j > 0; j--) {
        idx_t q, d;
        if (dl.shift[j] >= 0) {
            q = rem >> dl.shift[j];
            d = rem & (idx_t)(dl.extent[j] - 1);
        } else {
            q = rem / dl.extent[j];
            d = rem - q * dl.extent[j];
        }
        off += (int64_t)d * dl.stride[j];
        rem = q;
    }
    return off + (int64_t)rem * dl.stride[0];
}

template <class T, bool SMALL>
__global__ void __launch_bounds__(256) permute_gather_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                              const DigitList dl, int64_t total, double scale,
                                                              bool do_scale) {
    using idx_t = typename std::conditional<SMALL, uint32_t, uint64_t>::type;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        int64_t off = digit_offset<idx_t>(dl, (idx_t)o);
        T v = src[off];
        if (do_scale) v = (T)((double)v * scale);
        dst[o] = v;
    }
}

template <class T>
__global__ void __launch_bounds__(256) copy_scale_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                          int64_t total, double scale, bool do_scale) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        T v = src[o];
        if (do_scale) v = (T)((double)v * scale);
        dst[o] = v;
    }
}

template <class T>
static int permute_typed(ndmps_ctx* ctx, const ndmps_plan* plan, bool inverse, const T* src, T* dst, double scale) {
    int64_t total = plan->total;
    if (total == 0) return NDMPS_OK;
    bool do_scale = scale != 1.0;
    int64_t want = (total + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 32;
    int grid = (int)(want < cap ? want : cap);
    if (plan->identity) {
        if (src == dst && !do_scale) return NDMPS_OK;
        copy_scale_kernel<T><<<grid, 256, 0, ctx->stream>>>(src, dst, total, scale, do_scale);
        NDMPS_LAUNCH_CHECK(ctx);
        return NDMPS_OK;
    }
    NDMPS_REQUIRE(src != dst, "ndmps permute: in-place permutation is not supported");
    const DigitList& dl = inverse ? plan->dec : plan->enc;
    if (total < (int64_t(1) << 32))
        permute_gather_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(src, dst, dl, total, scale, do_scale);
    else
        permute_gather_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(src, dst, dl, total, scale, do_scale);
    NDMPS_LAUNCH_CHECK(ctx);
    return NDMPS_OK;
}

int permute(ndmps_ctx* ctx, const ndmps_plan* plan, bool inverse, const void* src, void* dst, int dtype, double scale) {
    StageScope sc(ctx, ST_PERMUTE);
    if (dtype == NDMPS_F32) return permute_typed<float>(ctx, plan, inverse, (const float*)src, (float*)dst, scale);
    return permute_typed<double>(ctx, plan, inverse, (const double*)src, (double*)dst, scale);
}

}  // namespace ndmps

using namespace ndmps;

extern "C" {

int ndmps_plan_create(int ndim, const int64_t* shape, int levels, const int64_t* factors, ndmps_plan_t** out) {
    NDMPS_REQUIRE(out && shape && factors, "ndmps_plan_create: NULL argument");
    NDMPS_REQUIRE(ndim >= 1 && ndim <= 8, "ndmps_plan_create: ndim %d outside 1..8", ndim);
    NDMPS_REQUIRE(levels >= 1 && levels * ndim <= NDMPS_MAX_DIGITS,
                  "ndmps_plan_create: levels*ndim = %d exceeds %d", levels * ndim, NDMPS_MAX_DIGITS);
    ndmps_plan* p = new ndmps_plan();
    memset(p, 0, sizeof(*p));
    p->ndim = ndim;
    p->levels = levels;
    p->total = 1;
    for (int a = 0; a < ndim; a++) {
        if (shape[a] <= 0) { delete p; set_error("ndmps_plan_create: non-positive extent"); return NDMPS_ERR_INVALID; }
        p->shape[a] = shape[a];
        p->total *= shape[a];
        int64_t prod = 1;
        for (int l = 0; l < levels; l++) {
            int64_t f = factors[l * ndim + a];
            if (f <= 0) { delete p; set_error("ndmps_plan_create: non-positive factor"); return NDMPS_ERR_INVALID; }
            p->factors[l * ndim + a] = f;
            prod *= f;
        }
        if (prod != shape[a]) {
            delete p;
            set_error("ndmps_plan_create: factors of axis %d multiply to %lld, extent is %lld", a, (long long)prod,
                      (long long)shape[a]);
            return NDMPS_ERR_INVALID;
        }
    }
    for (int l = 0; l < levels; l++) {
        int64_t d = 1;
        for (int a = 0; a < ndim; a++) d *= factors[l * ndim + a];
        p->site_dims[l] = d;
    }
    // strides of digit (l, a) in the volume (C order over axes, then over the axis' own digits)
    // and in the site-ordered array (C order over levels, then over axes inside a level)
    std::vector<int64_t> vol_stride(levels * ndim), site_stride(levels * ndim);
    int64_t axis_stride = 1;
    for (int a = ndim - 1; a >= 0; a--) {
        int64_t s = axis_stride;
        for (int l = levels - 1; l >= 0; l--) { vol_stride[l * ndim + a] = s; s *= factors[l * ndim + a]; }
        axis_stride *= shape[a];
    }
    int64_t s = 1;
    for (int l = levels - 1; l >= 0; l--)
        for (int a = ndim - 1; a >= 0; a--) { site_stride[l * ndim + a] = s; s *= factors[l * ndim + a]; }
    // encode: destination walks (l, a) in C order, gathers from the volume
    for (int l = 0; l < levels; l++)
        for (int a = 0; a < ndim; a++) push_digit(p->enc, (uint32_t)factors[l * ndim + a], vol_stride[l * ndim + a]);
    // decode: destination walks (a, l) in C order, gathers from the site-ordered array
    for (int a = 0; a < ndim; a++)
        for (int l = 0; l < levels; l++) push_digit(p->dec, (uint32_t)factors[l * ndim + a], site_stride[l * ndim + a]);
    finish(p->enc);
    finish(p->dec);
    p->identity = (p->enc.n == 1 && p->enc.stride[0] <= 1);
    *out = p;
    return NDMPS_OK;
}

int ndmps_plan_destroy(ndmps_plan_t* plan) {
    delete plan;
    return NDMPS_OK;
}

int ndmps_plan_site_dims(const ndmps_plan_t* plan, int64_t* dims_out) {
    NDMPS_REQUIRE(plan && dims_out, "ndmps_plan_site_dims: NULL argument");
    for (int l = 0; l < plan->levels; l++) dims_out[l] = plan->site_dims[l];
    return NDMPS_OK;
}

int ndmps_plan_debug_offsets(const ndmps_plan_t* plan, int inverse, int64_t first, int64_t count, int64_t* out_host) {
    NDMPS_REQUIRE(plan && out_host, "ndmps_plan_debug_offsets: NULL argument");
    NDMPS_REQUIRE(first >= 0 && count >= 0 && first + count <= plan->total, "ndmps_plan_debug_offsets: range outside the volume");
    const DigitList& dl = inverse ? plan->dec : plan->enc;
    for (int64_t i = 0; i < count; i++)
        out_host[i] = plan->identity ? first + i : digit_offset<uint64_t>(dl, (uint64_t)(first + i));
    return NDMPS_OK;
}

int ndmps_encode(ndmps_ctx_t* ctx, const ndmps_plan_t* plan, const void* src, void* dst, int dtype, double scale) {
    NDMPS_REQUIRE(ctx && plan && src && dst, "ndmps_encode: NULL argument");
    NDMPS_REQUIRE(dtype_ok(dtype), "ndmps_encode: bad dtype %d", dtype);
    return permute(ctx, plan, false, src, dst, dtype, scale);
}

int ndmps_decode(ndmps_ctx_t* ctx, const ndmps_plan_t* plan, const void* src, void* dst, int dtype) {
    NDMPS_REQUIRE(ctx && plan && src && dst, "ndmps_decode: NULL argument");
    NDMPS_REQUIRE(dtype_ok(dtype), "ndmps_decode: bad dtype %d", dtype);
    return permute(ctx, plan, true, src, dst, dtype, 1.0);
}

}  // extern "C"
