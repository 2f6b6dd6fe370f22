"""Oracle: axis factorisation and the N-D -> MPS-site index map (CPU, numpy).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Restates
``/root/reference/src/imgcompressionmps/utils/core.py`` (cited per function)
and the scatter/gather of ``core/ndmps.py:66-71,144-148`` as a permutation.
Written independently of the reference source: scalar loops and explicit
mixed-radix arithmetic instead of the reference's broadcast formulation.
"""
from __future__ import annotations

import numpy as np

INT64_MAX = np.iinfo(np.int64).max


def prime_factors(n: int) -> list[int]:
    """Ascending prime factors with multiplicity (reference: sympy ``factorint``
    at ``utils/core.py:107``; 1 -> [1] per ``utils/core.py:103-104``)."""
    if n == 1:
        return [1]
    out, p = [], 2
    while p * p <= n:
        while n % p == 0:
            out.append(p)
            n //= p
        p += 1 if p == 2 else 2
    if n > 1:
        out.append(n)
    return out


def _check_shape(shape):
    # utils/core.py:22-25 and :95-98
    if len(shape) == 0:
        raise ValueError("Shape cannot be empty.")
    for d in shape:
        if not isinstance(d, int) or isinstance(d, bool) or d <= 0:
            raise ValueError("All dimensions must be positive integers.")


def balance_factors(factors, target_num):
    """Merge the two smallest factors until ``target_num`` remain
    (``utils/core.py:38-76``)."""
    if target_num < 0:
        raise ValueError("target_num must be non-negative.")
    if target_num == 0 and len(factors) > 0:
        raise ValueError("Cannot reduce non-empty factor list to length zero.")
    work = sorted(factors)
    if len(work) < target_num:
        raise ValueError("The number of balanced factors cannot be less than the target number.")
    while len(work) > target_num:
        a = work.pop(0)
        b = work.pop(0)
        work.append(a * b)
        work.sort()
    return work


def get_factorlist(shape):
    """(L, ndim) per-level factors and (L+1, ndim) suffix products
    (``utils/core.py:79-126``): balance every axis to the shortest prime list,
    reverse odd-numbered axes, products row 0 = INT64_MAX, row L = 1."""
    _check_shape(shape)
    per_axis = [prime_factors(d) for d in shape]
    depth = min(len(f) for f in per_axis)
    per_axis = [balance_factors(f, depth) for f in per_axis]
    for a in range(1, len(per_axis), 2):
        per_axis[a] = per_axis[a][::-1]
    ndim = len(shape)
    fac = np.empty((depth, ndim), dtype=np.int64)
    for a in range(ndim):
        for lvl in range(depth):
            fac[lvl, a] = per_axis[a][lvl]
    prod = np.ones((depth + 1, ndim), dtype=np.int64)
    for a in range(ndim):
        running = 1
        for lvl in range(depth - 1, 0, -1):
            running *= int(fac[lvl, a])
            prod[lvl, a] = running
    prod[0, :] = INT64_MAX
    return fac, prod


def hierarchical_block_indexing(index, prod_block_sizes):
    """digit[l, a] = (x_a mod prod[l, a]) // prod[l+1, a]  (``utils/core.py:129-168``)."""
    index = np.asarray(index)
    prod_block_sizes = np.asarray(prod_block_sizes)
    ndim = index.shape[0]
    if prod_block_sizes.ndim != 2 or prod_block_sizes.shape[1] != ndim or prod_block_sizes.shape[0] < 2:
        raise ValueError("prod_block_sizes must be of shape (num_levels + 1, ndim) with ndim matching index.")
    levels = prod_block_sizes.shape[0] - 1
    out = np.empty((levels,) + index.shape, dtype=np.int64)
    for lvl in range(levels):
        for a in range(ndim):
            hi = int(prod_block_sizes[lvl, a])
            lo = int(prod_block_sizes[lvl + 1, a])
            out[lvl, a] = (index[a].astype(np.int64) % hi) // lo
    return out


def gen_encoding_map(shape):
    """Site dims (L,) and int64 map (L, *shape): site index of every voxel at
    every level, C-order over axes (``utils/core.py:6-35``)."""
    _check_shape(shape)
    fac, prod = get_factorlist(shape)
    digits = hierarchical_block_indexing(np.indices(shape), prod)
    levels, ndim = fac.shape
    enc = np.zeros((levels,) + tuple(shape), dtype=np.int64)
    for lvl in range(levels):
        for a in range(ndim):
            enc[lvl] = enc[lvl] * int(fac[lvl, a]) + digits[lvl, a]
    return np.prod(fac, axis=1), enc


def site_dims(shape):
    fac, _ = get_factorlist(shape)
    return [int(v) for v in np.prod(fac, axis=1)]


def encode(tensor):
    """The scatter of ``core/ndmps.py:66-71`` as a permutation: split axis a
    into its level digits, bring levels to the front, fuse the axes of a level.
    Returns the dense array shaped by the site dims."""
    tensor = np.asarray(tensor)
    fac, _ = get_factorlist(tuple(int(s) for s in tensor.shape))
    levels, ndim = fac.shape
    split = tensor.reshape([int(fac[lvl, a]) for a in range(ndim) for lvl in range(levels)])
    order = [a * levels + lvl for lvl in range(levels) for a in range(ndim)]
    return np.ascontiguousarray(split.transpose(order)).reshape([int(v) for v in np.prod(fac, axis=1)])


def decode(dense, shape):
    """Inverse of :func:`encode` (the gather of ``core/ndmps.py:144-148``)."""
    shape = tuple(int(s) for s in shape)
    fac, _ = get_factorlist(shape)
    levels, ndim = fac.shape
    split = np.asarray(dense).reshape([int(fac[lvl, a]) for lvl in range(levels) for a in range(ndim)])
    order = [lvl * ndim + a for a in range(ndim) for lvl in range(levels)]
    return np.ascontiguousarray(split.transpose(order)).reshape(shape)


def encode_by_map(tensor):
    """Literal scatter through the int64 map, as the reference does it
    (``core/ndmps.py:57-71``).  Slow; used to validate :func:`encode`."""
    tensor = np.asarray(tensor)
    dims, enc = gen_encoding_map(tuple(int(s) for s in tensor.shape))
    enc = np.moveaxis(enc, 0, -1).reshape(-1, len(dims))
    out = np.empty(tuple(int(d) for d in dims), dtype=tensor.dtype)
    out[tuple(enc[:, k] for k in range(len(dims)))] = tensor.reshape(-1)
    return out
