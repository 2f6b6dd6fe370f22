"""Host-side planning of the N-D -> MPS-site encoding.

Mirrors the public functions of the reference's ``utils/core.py`` (same names,
arguments, return values and ``ValueError`` behaviour):

* ``balance_factors``              <- ``utils/core.py:38-76``
* ``get_factorlist``               <- ``utils/core.py:79-126``
* ``hierarchical_block_indexing``  <- ``utils/core.py:129-168``
* ``gen_encoding_map``             <- ``utils/core.py:6-35``

This is integer bookkeeping on a handful of numbers per axis; it stays on the host.
The expensive part of the reference - materialising the int64 map (8*L bytes per
voxel) and scattering through it - is replaced on the device by the permutation
kernels (``ndmps_encode`` / ``ndmps_decode``), which consume only the (L, ndim)
factor table produced here.  ``gen_encoding_map`` still builds the full map for
callers that ask for it (kept for API compatibility; memory-lean integer version).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

_INT64_MAX = int(np.iinfo(np.int64).max)


def _validate_shape(shape) -> None:
    if len(shape) == 0:
        raise ValueError("Shape cannot be empty.")
    if any(not isinstance(dim, int) or isinstance(dim, bool) or dim <= 0 for dim in shape):
        raise ValueError("All dimensions must be positive integers.")


def _prime_factors(n: int) -> List[int]:
    """Prime factors with multiplicity, ascending (trial division; axes are small)."""
    found: List[int] = []
    while n % 2 == 0:
        found.append(2)
        n //= 2
    f = 3
    while f * f <= n:
        if n % f == 0:
            found.append(f)
            n //= f
        else:
            f += 2
    if n > 1:
        found.append(n)
    return found


def balance_factors(factors: List[int], target_num: int) -> List[int]:
    """Fuse the two smallest factors until ``target_num`` factors remain; sorted result."""
    if target_num < 0:
        raise ValueError("target_num must be non-negative.")
    if target_num == 0 and len(factors) > 0:
        raise ValueError("Cannot reduce non-empty factor list to length zero.")
    merged = sorted(factors)
    if len(merged) < target_num:
        raise ValueError("The number of balanced factors cannot be less than the target number.")
    while len(merged) > target_num:
        merged = sorted([merged[0] * merged[1]] + merged[2:])
    return merged


def get_factorlist(shape: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """Balanced per-level factors ``(L, ndim)`` and suffix products ``(L+1, ndim)``.

    Every axis is prime-factorised, balanced to the shortest factor list, and every
    second axis is reversed (the reference's "snake").  Products: row 0 is INT64_MAX,
    row l (1 <= l < L) is the product of the axis' factors at levels >= l, row L is 1.
    """
    _validate_shape(shape)
    lists = [[1] if dim == 1 else _prime_factors(dim) for dim in shape]
    depth = min(len(f) for f in lists)
    lists = [balance_factors(f, depth) for f in lists]
    lists = [f[::-1] if axis % 2 == 1 else f for axis, f in enumerate(lists)]
    factor_arr = np.array(lists, dtype=np.int64).T.copy()
    suffix = np.ones((depth + 1, len(shape)), dtype=np.int64)
    for lvl in range(depth - 1, 0, -1):
        suffix[lvl] = suffix[lvl + 1] * factor_arr[lvl]
    suffix[0] = _INT64_MAX
    return factor_arr, suffix


def hierarchical_block_indexing(index: np.ndarray, prod_block_sizes: np.ndarray) -> np.ndarray:
    """Mixed-radix digits ``(L, ndim, *shape)`` of an index grid (integer arithmetic)."""
    index = np.asarray(index)
    prod_block_sizes = np.asarray(prod_block_sizes)
    ndim = index.shape[0]
    if prod_block_sizes.ndim != 2 or prod_block_sizes.shape[1] != ndim or prod_block_sizes.shape[0] < 2:
        raise ValueError("prod_block_sizes must be of shape (num_levels + 1, ndim) with ndim matching index.")
    bcast = (slice(None), slice(None)) + (None,) * (index.ndim - 1)
    upper = prod_block_sizes[:-1].astype(np.int64)[bcast]
    lower = prod_block_sizes[1:].astype(np.int64)[bcast]
    return (index[None].astype(np.int64) % upper) // lower


def gen_encoding_map(shape: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """Site dimensions ``(L,)`` and the int64 map ``(L, *shape)`` voxel -> site index."""
    _validate_shape(shape)
    factors, _ = get_factorlist(shape)
    levels, ndim = factors.shape
    enc = np.zeros((levels,) + tuple(shape), dtype=np.int64)
    # digit of axis a at level l depends on the axis coordinate only: build it per axis
    # and broadcast, instead of the (L, ndim, *shape) temporaries of the reference
    for a in range(ndim):
        coord = np.arange(shape[a], dtype=np.int64)
        view = [1] * ndim
        view[a] = shape[a]
        weight_below = 1
        digits = [None] * levels
        for lvl in range(levels - 1, -1, -1):
            digits[lvl] = (coord // weight_below) % int(factors[lvl, a])
            weight_below *= int(factors[lvl, a])
        for lvl in range(levels):
            enc[lvl] = enc[lvl] * int(factors[lvl, a]) + digits[lvl].reshape(view)
    return np.prod(factors, axis=1), enc
