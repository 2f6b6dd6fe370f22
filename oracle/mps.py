"""Oracle: the matrix-product-state algebra the reference delegates to quimb 1.9.0.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  **Parity unpinned**: quimb
is a third-party dependency of the reference (``requirements.txt:110``), not
vendored and not installable here; this file restates its published algorithm
 at the reference's call sites:

* ``tt_svd``            <- ``qtn.MatrixProductState.from_dense`` (``core/ndmps.py:74``)
* ``compress_bond``     <- ``qtn.tensor_compress_bond(..., cutoff_mode="rel")`` (``core/ndmps.py:104-106``)
* ``overlap``           <- ``mps @ mps`` (``core/ndmps.py:76,86``, ``utils/metrics.py:160``)
* ``contract_dense``    <- ``mps ^ ...`` + ``moveindex`` (``core/ndmps.py:140-142``)

Core layout everywhere: site 0 ``(d0, r0)``, middle ``(r_{i-1}, d_i, r_i)``,
last ``(r_{L-2}, d_{L-1})``, C-order, float64.  A one-site MPS is a single
``(d0,)`` vector.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# singular-value trimming (quimb.tensor.decomp: number-to-keep + renorm factor)
# ----------------------------------------------------------------------------
def n_keep(s, cutoff, cutoff_mode, max_bond=None):
    """How many singular values survive.

    ``rel``  : keep ``s > cutoff * s[0]``.
    ``rsum2``: drop the longest tail whose sum of squares is <= cutoff * sum of
    all squares (scan from the smallest value; stop at the first value pushing
    the running sum strictly above the target).
    ``abs``  : keep ``s > cutoff``.
    A cutoff of exactly 0 disables cutoff-trimming (only ``max_bond`` applies).
    Always at least one value is kept.
    """
    s = np.asarray(s, dtype=np.float64)
    n = s.size
    if cutoff > 0.0:
        if cutoff_mode == "abs":
            n = int(np.sum(s > cutoff))
        elif cutoff_mode == "rel":
            n = int(np.sum(s > cutoff * s[0]))
        elif cutoff_mode in ("rsum2", "sum2", "rsum1", "sum1"):
            power = 2 if cutoff_mode.endswith("2") else 1
            sp = s ** power
            target = cutoff * (np.nansum(sp) if cutoff_mode.startswith("r") else 1.0)
            run = 0.0
            n = s.size
            for i in range(s.size - 1, -1, -1):
                if not np.isnan(sp[i]):
                    run += sp[i]
                if run > target:
                    break
                n -= 1
        else:
            raise ValueError(f"unknown cutoff_mode {cutoff_mode!r}")
        n = max(n, 1)
    if max_bond is not None and max_bond > 0:
        n = min(n, int(max_bond))
    return n


def renorm_factor(s, n, power):
    """((kept + lost) / kept) ** (1/power) over s**power; 1.0 if nothing lost."""
    if n >= len(s) or power <= 0:
        return 1.0
    sp = np.asarray(s, dtype=np.float64) ** power
    keep = np.nansum(sp[:n])
    lose = np.nansum(sp[n:])
    return float(((keep + lose) / keep) ** (1.0 / power))


def _svd(m):
    try:
        return np.linalg.svd(m, full_matrices=False)          # LAPACK gesdd, as quimb's numba path
    except np.linalg.LinAlgError:                               # quimb falls back to gesvd
        import scipy.linalg
        return scipy.linalg.svd(m, full_matrices=False, lapack_driver="gesvd")


# ----------------------------------------------------------------------------
# from_dense: left -> right TT-SVD, weight absorbed to the right
# ----------------------------------------------------------------------------
def tt_svd(dense, dims, cutoff=1e-10, cutoff_mode="rsum2", max_bond=None, renorm=None,
           return_svals=False):
    """Left-canonical MPS of ``dense`` reshaped to ``dims`` (Appendix A.1).

    Defaults are those ``from_dense`` applies when the reference calls it with
    no options: ``rsum2`` cutoff 1e-10, renorm power 2 (auto for rsum2), singular
    values absorbed right so all weight ends in the last core.
    """
    dims = [int(d) for d in dims]
    L = len(dims)
    if renorm is None:
        renorm = {"rsum2": 2, "sum2": 2, "rsum1": 1, "sum1": 1}.get(cutoff_mode, 0)
    tm = np.asarray(dense, dtype=np.float64).reshape(dims)
    if L == 1:
        return ([tm.copy()], []) if return_svals else [tm.copy()]
    cores, svals = [], []
    r_prev = 1
    tm = tm.reshape(dims[0], -1)
    for i in range(L - 1):
        m = tm.reshape(r_prev * dims[i], -1)
        u, s, vh = _svd(m)
        n = n_keep(s, cutoff, cutoff_mode, max_bond)
        f = renorm_factor(s, n, renorm)
        s_kept = s[:n] * f
        svals.append(s_kept.copy())
        u = u[:, :n]
        cores.append(u.reshape(dims[0], n) if i == 0 else u.reshape(r_prev, dims[i], n))
        tm = s_kept[:, None] * vh[:n]
        r_prev = n
    cores.append(tm.reshape(r_prev, dims[-1]))
    return (cores, svals) if return_svals else cores


# ----------------------------------------------------------------------------
# tensor_compress_bond, cutoff_mode="rel", absorb="both", reduced=True
# ----------------------------------------------------------------------------
def _as_left_matrix(core, first):
    return core if first else core.reshape(core.shape[0] * core.shape[1], core.shape[2])


def compress_bond(t1, t2, first, last, cutoff, cutoff_mode="rel", max_bond=None, renorm=0):
    """Truncate the bond between neighbouring cores (Appendix A.2).

    QR of the left core over its non-shared indices, LQ of the right core,
    SVD of R.L, keep by the ``rel`` rule (no renormalisation), split sqrt(s) to
    both sides.  Returns the two new cores and the kept singular values.
    """
    a = _as_left_matrix(t1, first)                                  # (rows, r)
    b = t2 if last else t2.reshape(t2.shape[0], t2.shape[1] * t2.shape[2])   # (r, cols)
    q1, r = np.linalg.qr(a)
    q2t, lt = np.linalg.qr(b.T)
    u, s, vh = _svd(r @ lt.T)
    n = n_keep(s, cutoff, cutoff_mode, max_bond)
    f = renorm_factor(s, n, renorm)
    s_kept = s[:n] * f
    root = np.sqrt(s_kept)
    a_new = q1 @ (u[:, :n] * root[None, :])
    b_new = (root[:, None] * vh[:n]) @ q2t.T
    t1_new = a_new if first else a_new.reshape(t1.shape[0], t1.shape[1], n)
    t2_new = b_new if last else b_new.reshape(n, t2.shape[1], t2.shape[2])
    return t1_new, t2_new, s_kept


def compress_all(cores, cutoff, max_bond=None):
    """The loop of ``NDMPS.compress`` (``core/ndmps.py:103-106``): bonds left to
    right, no canonicalisation in between.  Returns new core list + svals."""
    cores = [np.array(c, dtype=np.float64) for c in cores]
    svals = []
    L = len(cores)
    for i in range(1, L):
        cores[i - 1], cores[i], s = compress_bond(cores[i - 1], cores[i], first=(i == 1), last=(i == L - 1),
                                                  cutoff=cutoff, max_bond=max_bond)
        svals.append(s)
    return cores, svals


# ----------------------------------------------------------------------------
# contractions
# ----------------------------------------------------------------------------
def contract_dense(cores):
    """Cumulative left -> right contraction to the dense site-ordered tensor (A.3)."""
    L = len(cores)
    if L == 1:
        return np.array(cores[0], dtype=np.float64)
    dims = [cores[0].shape[0]] + [c.shape[1] for c in cores[1:]]
    x = np.asarray(cores[0], dtype=np.float64)                       # (d0, r0)
    for k in range(1, L - 1):
        c = cores[k]
        x = (x @ c.reshape(c.shape[0], -1)).reshape(-1, c.shape[2])
    x = x @ cores[-1]
    return x.reshape(dims)


def overlap(a, b):
    """sum over all entries of dense(a) * dense(b) via transfer matrices (A.3);
    no conjugation (real data)."""
    L = len(a)
    if L == 1:
        return float(np.dot(np.ravel(a[0]), np.ravel(b[0])))
    e = a[0].T @ b[0]                                                # (ra, rb)
    for k in range(1, L - 1):
        t = np.tensordot(e, b[k], axes=(1, 0))                       # (ra, d, rb')
        e = np.tensordot(a[k], t, axes=([0, 1], [0, 1]))             # (ra', rb')
    return float(np.sum((a[-1] @ b[-1].T) * e))


def bond_sizes(cores):
    if len(cores) == 1:
        return []
    return [int(cores[0].shape[1])] + [int(c.shape[2]) for c in cores[1:-1]]


def num_elements(cores):
    return int(sum(int(c.size) for c in cores))
