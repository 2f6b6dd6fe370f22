"""Thin torch-tensor wrappers over the C ABI (``include/ndmps.h``).

Every function here launches this repo's own CUDA kernels on torch's current
stream; torch only provides the device buffers.  Nothing falls back to torch math.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import numpy as np

from . import _native as N
from .utils.core import get_factorlist


def _torch():
    import torch
    return torch


def _dev(t, what="tensor"):
    if not t.is_cuda:
        raise ValueError(f"{what} must live on the CUDA device")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    return t


@lru_cache(maxsize=64)
def plan_for(shape) -> N.Plan:
    shape = tuple(int(s) for s in shape)
    factors, _ = get_factorlist(shape)
    return N.Plan(shape, factors)


# ---- K1 ---------------------------------------------------------------------------
def encode(volume, scale: float = 1.0, plan=None):
    """volume (original shape, device) -> dense array shaped by the site dims."""
    torch = _torch()
    plan = plan or plan_for(tuple(volume.shape))
    _dev(volume, "volume")
    out = torch.empty(plan.site_dims, dtype=volume.dtype, device=volume.device)
    N.check(N.load_library().ndmps_encode(N.handle(), plan.handle, N.ptr(volume), N.ptr(out),
                                          N.dtype_code(volume.dtype), float(scale)), "ndmps_encode")
    return out


def decode(dense, shape, plan=None):
    torch = _torch()
    plan = plan or plan_for(tuple(shape))
    _dev(dense, "dense")
    if dense.numel() != plan.total:
        raise ValueError("dense array does not match the volume shape")
    out = torch.empty(plan.shape, dtype=dense.dtype, device=dense.device)
    N.check(N.load_library().ndmps_decode(N.handle(), plan.handle, N.ptr(dense), N.ptr(out),
                                          N.dtype_code(dense.dtype)), "ndmps_decode")
    return out


# ---- reductions ---------------------------------------------------------------------
def sumsq(x) -> float:
    _dev(x)
    out = C.c_double()
    N.check(N.load_library().ndmps_sumsq(N.handle(), N.ptr(x), x.numel(), N.dtype_code(x.dtype), C.byref(out)),
            "ndmps_sumsq")
    return out.value


def minmax(tensors) -> np.ndarray:
    tensors = [_dev(t) for t in tensors]
    if not tensors:
        return np.zeros((0, 2))
    code = N.dtype_code(tensors[0].dtype)
    out = np.empty((len(tensors), 2), dtype=np.float64)
    N.check(N.load_library().ndmps_minmax(N.handle(), N.ptr_array(tensors), N.i64_array([t.numel() for t in tensors]),
                                          len(tensors), code, out.ctypes.data_as(N.p_f64)), "ndmps_minmax")
    return out


def psnr_terms(a, b):
    _dev(a), _dev(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("PSNR operands must share shape and dtype")
    out = (C.c_double * 2)()
    N.check(N.load_library().ndmps_psnr_terms(N.handle(), N.ptr(a), N.ptr(b), a.numel(), N.dtype_code(a.dtype), out),
            "ndmps_psnr_terms")
    return out[0], out[1]


# ---- K11 ---------------------------------------------------------------------------
def dct_last_axis(x, inverse: bool = False):
    torch = _torch()
    _dev(x)
    n = int(x.shape[-1])
    out = torch.empty_like(x)
    N.check(N.load_library().ndmps_dct_last_axis(N.handle(), N.ptr(x), N.ptr(out), x.numel() // n, n,
                                                 int(bool(inverse)), N.dtype_code(x.dtype)), "ndmps_dct_last_axis")
    return out


# ---- building blocks ---------------------------------------------------------------
def gram(m, side: int = 0):
    torch = _torch()
    _dev(m)
    rows, cols = int(m.shape[0]), int(m.shape[1])
    n = rows if side == 0 else cols
    g = torch.empty((n, n), dtype=torch.float64, device=m.device)
    N.check(N.load_library().ndmps_gram(N.handle(), N.ptr(m), rows, cols, cols, N.dtype_code(m.dtype), int(side),
                                        N.ptr(g)), "ndmps_gram")
    return g


def eigh(a):
    """Eigen-decomposition of a symmetric float64 matrix (descending); returns (evals, evecs, sweeps)."""
    torch = _torch()
    _dev(a)
    if a.dtype != torch.float64 or a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("eigh expects a square float64 matrix")
    n = int(a.shape[0])
    work = a.clone()
    evals = torch.empty(n, dtype=torch.float64, device=a.device)
    evecs = torch.empty((n, n), dtype=torch.float64, device=a.device)
    sweeps = C.c_int(0)
    N.check(N.load_library().ndmps_eigh(N.handle(), N.ptr(work), n, N.ptr(evals), N.ptr(evecs), C.byref(sweeps)),
            "ndmps_eigh")
    return evals, evecs, sweeps.value


def eigh_topk(a, k: int):
    """Leading k eigenpairs of a symmetric PSD float64 matrix; returns (evals[k] descending, evecs[n, k], trace, health)."""
    torch = _torch()
    _dev(a)
    if a.dtype != torch.float64 or a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("eigh_topk expects a square float64 matrix")
    n, k = int(a.shape[0]), int(k)
    a = a.contiguous()
    out = torch.empty(k + 2, dtype=torch.float64, device=a.device)
    evecs = torch.empty((n, k), dtype=torch.float64, device=a.device)
    N.check(N.load_library().ndmps_eigh_topk(N.handle(), N.ptr(a), n, k, N.ptr(out), N.ptr(evecs)), "ndmps_eigh_topk")
    host = out.cpu()
    return out[:k], evecs, float(host[k]), int(host[k + 1])


def gemm(a, b, out_dtype=None, alpha: float = 1.0):
    """a @ b for 2-D (possibly transposed-view) device tensors through ndmps_gemm."""
    torch = _torch()
    m, k = a.shape
    k2, n = b.shape
    if k != k2:
        raise ValueError("gemm: inner dimensions differ")
    out_dtype = out_dtype or a.dtype
    c = torch.empty((m, n), dtype=out_dtype, device=a.device)
    N.check(N.load_library().ndmps_gemm(N.handle(), m, n, k, float(alpha),
                                        N.ptr(a), N.dtype_code(a.dtype), a.stride(0), a.stride(1),
                                        N.ptr(b), N.dtype_code(b.dtype), b.stride(0), b.stride(1),
                                        N.ptr(c), N.dtype_code(out_dtype), n), "ndmps_gemm")
    return c


# ---- sweep / truncation / contraction -------------------------------------------------
def bond_bounds(dims, max_bond=None):
    """Upper bound of every bond: min(prod left, prod right, max_bond)."""
    L = len(dims)
    out = []
    left = 1
    for i in range(L - 1):
        left *= int(dims[i])
        right = 1
        for d in dims[i + 1:]:
            right *= int(d)
        bound = min(left, right)
        # a bond can also never exceed (previous bond * d_i)
        if out:
            bound = min(bound, out[-1] * int(dims[i]))
        if max_bond:
            bound = min(bound, int(max_bond))
        out.append(bound)
    return out


def core_shape(i, L, dims, ranks):
    if L == 1:
        return (int(dims[0]),)
    if i == 0:
        return (int(dims[0]), int(ranks[0]))
    if i == L - 1:
        return (int(ranks[L - 2]), int(dims[L - 1]))
    return (int(ranks[i - 1]), int(dims[i]), int(ranks[i]))


def ttsvd(dense, dims, cutoff=1e-10, cutoff_mode="rsum2", max_bond=None, renorm=None):
    """Left-canonical MPS cores of ``dense`` (site order); returns (cores, ranks, svals)."""
    torch = _torch()
    _dev(dense, "dense")
    dims = [int(d) for d in dims]
    L = len(dims)
    mode = N.CUTOFF_MODES[cutoff_mode]
    if renorm is None:
        renorm = {N.CUT_RSUM2: 2, N.CUT_SUM2: 2, N.CUT_RSUM1: 1, N.CUT_SUM1: 1}.get(mode, 0)
    bounds = bond_bounds(dims, max_bond)
    caps = [int(np.prod(core_shape(i, L, dims, bounds), dtype=np.int64)) for i in range(L)]
    bufs = [torch.empty(c, dtype=dense.dtype, device=dense.device) for c in caps]
    ranks = (C.c_int64 * max(L - 1, 1))()
    stride = max(bounds) if bounds else 1
    svals = np.zeros((max(L - 1, 1), stride), dtype=np.float64)
    N.check(N.load_library().ndmps_ttsvd(N.handle(), N.ptr(dense), N.dtype_code(dense.dtype), L, N.i64_array(dims),
                                         float(cutoff), mode, int(max_bond or 0), int(renorm),
                                         N.ptr_array(bufs), N.i64_array(caps), ranks,
                                         svals.ctypes.data_as(N.p_f64), stride), "ndmps_ttsvd")
    ranks = [int(r) for r in ranks][:L - 1]
    cores = []
    for i in range(L):
        shp = core_shape(i, L, dims, ranks)
        n = int(np.prod(shp, dtype=np.int64))
        cores.append(bufs[i][:n].view(shp) if n == caps[i] else bufs[i][:n].clone().view(shp))
    return cores, ranks, [svals[i, :ranks[i]].copy() for i in range(L - 1)]


class _DevicePointer:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, count: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def ttsvd_sharded(dense_local, dims_local, world: int, allreduce=None, stop_bytes: int = 8 << 20, cutoff=1e-10,
                  cutoff_mode="rsum2", max_bond=None, renorm=None):
    """Sharded phase of the sweep on this rank's column block (``ndmps_ttsvd_sharded``).

    ``allreduce(tensor)`` must sum the float64 device tensor in place over the ranks (e.g.
    ``torch.distributed.all_reduce``); it is called once per Gram pass on the current stream.
    Returns (cores of the finished sites, their ranks, singular values, remainder[rows, cols])."""
    torch = _torch()
    _dev(dense_local, "dense_local")
    dims = [int(d) for d in dims_local]
    L = len(dims)
    mode = N.CUTOFF_MODES[cutoff_mode]
    if renorm is None:
        renorm = {N.CUT_RSUM2: 2, N.CUT_SUM2: 2, N.CUT_RSUM1: 1, N.CUT_SUM1: 1}.get(mode, 0)
    gdims = dims[:-1] + [dims[-1] * int(world)]
    bounds = bond_bounds(gdims, max_bond)
    caps = [int(np.prod(core_shape(i, L, gdims, bounds), dtype=np.int64)) for i in range(L)]
    bufs = [torch.empty(c, dtype=dense_local.dtype, device=dense_local.device) for c in caps]
    ranks = (C.c_int64 * max(L - 1, 1))()
    stride = max(bounds) if bounds else 1
    svals = np.zeros((max(L - 1, 1), stride), dtype=np.float64)
    remainder = torch.empty(dense_local.numel(), dtype=dense_local.dtype, device=dense_local.device)
    sites_done = C.c_int(0)
    rshape = (C.c_int64 * 2)()
    failure = []

    def hook(_user, buf, count, _stream):
        try:
            allreduce(torch.as_tensor(_DevicePointer(buf, count), device=dense_local.device))
            return 0
        except Exception as exc:                     # never let an exception cross the C frame
            failure.append(exc)
            return 1

    cb = N.ALLREDUCE_FN(hook) if (world > 1 and allreduce is not None) else C.cast(None, N.ALLREDUCE_FN)
    if world > 1 and allreduce is None:
        raise ValueError("ttsvd_sharded: world > 1 needs an allreduce callable")
    rc = N.load_library().ndmps_ttsvd_sharded(
        N.handle(), N.ptr(dense_local), N.dtype_code(dense_local.dtype), L, N.i64_array(dims), int(world), cb, None,
        int(stop_bytes), float(cutoff), mode, int(max_bond or 0), int(renorm), N.ptr_array(bufs), N.i64_array(caps), ranks,
        svals.ctypes.data_as(N.p_f64), stride, C.byref(sites_done), N.ptr(remainder), remainder.numel(), rshape)
    if failure:
        raise failure[0]
    N.check(rc, "ndmps_ttsvd_sharded")
    done = sites_done.value
    rk = [int(r) for r in ranks][:done]
    cores = []
    for i in range(done):                              # done < L - 1: never the last core
        shp = (gdims[0], rk[0]) if i == 0 else (rk[i - 1], gdims[i], rk[i])
        n = int(np.prod(shp, dtype=np.int64))
        cores.append(bufs[i][:n].clone().view(shp))
    rows, cols = int(rshape[0]), int(rshape[1])
    return cores, rk, [svals[i, :rk[i]].copy() for i in range(done)], remainder[:rows * cols].view(rows, cols)


def interleave_shards(gathered, world: int, rows: int, cmid: int, dl: int):
    """gathered[g][r][c][t] (rank-major, as all-gathered) -> out[r][c][g][t] (site order)."""
    torch = _torch()
    _dev(gathered, "gathered")
    out = torch.empty(rows * cmid * world * dl, dtype=gathered.dtype, device=gathered.device)
    N.check(N.load_library().ndmps_interleave_shards(N.handle(), N.ptr(gathered), N.dtype_code(gathered.dtype), int(world),
                                                     int(rows), int(cmid), int(dl), N.ptr(out)), "ndmps_interleave_shards")
    return out.view(rows, cmid * world * dl)


def compress_bond(t1, t2, cutoff, cutoff_mode="rel", max_bond=None, renorm=0):
    """Truncate the bond between 2-D matricised cores t1 (a x r) and t2 (r x b)."""
    torch = _torch()
    _dev(t1), _dev(t2)
    a, r = int(t1.shape[0]), int(t1.shape[1])
    r2, b = int(t2.shape[0]), int(t2.shape[1])
    if r != r2 or t1.dtype != t2.dtype:
        raise ValueError("compress_bond: cores do not share the bond")
    o1 = torch.empty(a * r, dtype=t1.dtype, device=t1.device)
    o2 = torch.empty(r * b, dtype=t1.dtype, device=t1.device)
    n = C.c_int64(0)
    svals = np.zeros(r, dtype=np.float64)
    N.check(N.load_library().ndmps_compress_bond(N.handle(), N.ptr(t1), N.ptr(t2), N.dtype_code(t1.dtype), a, r, b,
                                                 float(cutoff), N.CUTOFF_MODES[cutoff_mode], int(max_bond or 0),
                                                 int(renorm), N.ptr(o1), N.ptr(o2), C.byref(n),
                                                 svals.ctypes.data_as(N.p_f64)), "ndmps_compress_bond")
    n = int(n.value)
    return o1[:a * n].clone().view(a, n), o2[:n * b].clone().view(n, b), svals[:n].copy()


def _ranks_of(cores):
    L = len(cores)
    if L == 1:
        return []
    return [int(cores[0].shape[1])] + [int(c.shape[2]) for c in cores[1:-1]]


def _dims_of(cores):
    if len(cores) == 1:
        return [int(cores[0].shape[0])]
    return [int(cores[0].shape[0])] + [int(c.shape[1]) for c in cores[1:]]


def contract_dense(cores):
    torch = _torch()
    cores = [_dev(c, "core") for c in cores]
    dims, ranks = _dims_of(cores), _ranks_of(cores)
    out = torch.empty(dims, dtype=cores[0].dtype, device=cores[0].device)
    N.check(N.load_library().ndmps_contract_dense(N.handle(), N.ptr_array(cores), N.dtype_code(cores[0].dtype),
                                                  len(cores), N.i64_array(dims), N.i64_array(ranks), N.ptr(out)),
            "ndmps_contract_dense")
    return out


def overlap(cores_a, cores_b) -> float:
    cores_a = [_dev(c, "core") for c in cores_a]
    cores_b = [_dev(c, "core") for c in cores_b]
    dims = _dims_of(cores_a)
    if dims != _dims_of(cores_b):
        raise ValueError("overlap: the two MPS have different site dimensions")
    out = C.c_double()
    N.check(N.load_library().ndmps_overlap(N.handle(), N.ptr_array(cores_a), N.i64_array(_ranks_of(cores_a)),
                                           N.dtype_code(cores_a[0].dtype), N.ptr_array(cores_b),
                                           N.i64_array(_ranks_of(cores_b)), N.dtype_code(cores_b[0].dtype),
                                           len(cores_a), N.i64_array(dims), C.byref(out)), "ndmps_overlap")
    return out.value


# ---- K12 ---------------------------------------------------------------------------
def int_code(np_dtype) -> int:
    """The C ABI's name of an integer type: its bit count, negative when signed.  ValueError for anything
    ``np.iinfo`` rejects, as ``scale_to_dtype`` (``utils/filetools.py:26``) raises."""
    info = np.iinfo(np.dtype(np_dtype))            # ValueError for non-integer dtypes
    return int(info.bits) if info.min == 0 else -int(info.bits)


def _torch_int(code: int):
    torch = _torch()
    return {8: torch.uint8, 16: torch.uint16, 32: torch.uint32, 64: torch.uint64,
            -8: torch.int8, -16: torch.int16, -32: torch.int32, -64: torch.int64}[code]


def quantize(x, lo: float, hi: float, bits: int):
    """Min-max quantisation of a device tensor to the integer type `bits` names (8 / 16 / 32 / 64 unsigned, negative:
    signed) -- ``scale_to_dtype`` for every dtype ``np.iinfo`` knows."""
    torch = _torch()
    _dev(x)
    q = torch.empty(x.shape, dtype=_torch_int(int(bits)), device=x.device)
    N.check(N.load_library().ndmps_quantize(N.handle(), N.ptr(x), x.numel(), N.dtype_code(x.dtype), float(lo), float(hi),
                                            int(bits), N.ptr(q)), "ndmps_quantize")
    return q


def dequantize(q, lo: float, hi: float, bits: int, dtype):
    torch = _torch()
    _dev(q)
    x = torch.empty(q.shape, dtype=dtype, device=q.device)
    N.check(N.load_library().ndmps_dequantize(N.handle(), N.ptr(q), q.numel(), int(bits), float(lo), float(hi),
                                              N.dtype_code(dtype), N.ptr(x)), "ndmps_dequantize")
    return x


def to_host_int(q, np_dtype) -> np.ndarray:
    """Device integer tensor -> numpy array of `np_dtype` (bit reinterpretation where torch lacks numpy interop)."""
    np_dtype = np.dtype(np_dtype)
    try:
        out = q.cpu().numpy()
        if out.dtype == np_dtype:
            return out
    except (TypeError, RuntimeError):
        pass
    torch = _torch()
    raw = q.contiguous().view(torch.uint8).cpu().numpy()
    return raw.view(np_dtype).reshape(tuple(q.shape))


# ---- K9 ---------------------------------------------------------------------------
def ssim(a, b) -> float:
    _dev(a), _dev(b)
    out = C.c_double()
    N.check(N.load_library().ndmps_ssim(N.handle(), N.ptr(a), N.ptr(b), N.dtype_code(a.dtype), a.ndim,
                                        N.i64_array(a.shape), C.byref(out)), "ndmps_ssim")
    return out.value


def ssim_slices(a, b, axis: int) -> np.ndarray:
    _dev(a), _dev(b)
    out = np.empty(int(a.shape[axis]), dtype=np.float64)
    N.check(N.load_library().ndmps_ssim_slices(N.handle(), N.ptr(a), N.ptr(b), N.dtype_code(a.dtype),
                                               N.i64_array(a.shape), int(axis), out.ctypes.data_as(N.p_f64)),
            "ndmps_ssim_slices")
    return out


# ---- host-buffer end-to-end entry ---------------------------------------------------
def roundtrip_host(src: np.ndarray, max_bond=None, cutoff=1e-10, cutoff_mode="rsum2", renorm=None, out=None, extras=None):
    """NDMPS.from_tensor(src, max_bond=...).to_tensor() in ONE C-ABI call on host buffers
    (H2D, encode, sweep, boundary list + norm, contract, decode, D2H).  Returns (reconstruction, ranks);
    a dict passed as ``extras`` receives ``norm`` and ``boundary_list``."""
    if src.dtype not in (np.float32, np.float64):
        raise ValueError("roundtrip_host expects a float32 or float64 array")
    src = np.ascontiguousarray(src)
    plan = plan_for(tuple(src.shape))
    mode = N.CUTOFF_MODES[cutoff_mode]
    if renorm is None:
        renorm = {N.CUT_RSUM2: 2, N.CUT_SUM2: 2, N.CUT_RSUM1: 1, N.CUT_SUM1: 1}.get(mode, 0)
    if out is None:
        out = np.empty_like(src)
    ranks = (C.c_int64 * max(plan.levels - 1, 1))()
    code = N.F32 if src.dtype == np.float32 else N.F64
    norm = C.c_double(0.0)
    bounds = np.zeros((plan.levels, 2), dtype=np.float64)
    N.check(N.load_library().ndmps_roundtrip_host(N.handle(), plan.handle, src.ctypes.data_as(C.c_void_p),
                                                  out.ctypes.data_as(C.c_void_p), code, float(cutoff), mode,
                                                  int(max_bond or 0), int(renorm), ranks, C.byref(norm),
                                                  bounds.ctypes.data_as(N.p_f64)), "ndmps_roundtrip_host")
    if extras is not None:
        extras["norm"] = norm.value
        extras["boundary_list"] = bounds
    return out, [int(r) for r in ranks][:plan.levels - 1]
