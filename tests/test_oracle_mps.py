"""Oracle TT-SVD / bond compression / contractions: the behavioural invariants of the
reference's tests/core/test_ndmps.py at its seed and shapes, and the probe table of
SURVEY.md Appendix A.4 (the only numeric evidence tying the quimb restatement to the
reference).  Parity with quimb itself is UNPINNED (quimb is not installable here)."""
import copy
import math

import numpy as np
import pytest

from oracle import mps as M
from oracle.ndmps import OracleNDMPS


@pytest.fixture(scope="module")
def tensors():
    rng = np.random.default_rng(2025)                    # tests/core/test_ndmps.py:6-8
    return {"2d": rng.random((512, 680)), "3d": rng.random((8, 512, 680))}


def test_trim_rules():
    s = np.array([10.0, 5.0, 1.0, 1e-3, 1e-6])
    assert M.n_keep(s, 0.2, "rel") == 2                   # 5 > 2, 1 > 2 is false
    assert M.n_keep(s, 0.1, "rel") == 2                   # strict: 1.0 > 1.0 is false
    assert M.n_keep(s, 0.0, "rel") == 5                   # cutoff 0 disables trimming
    assert M.n_keep(s, 1e-10, "rsum2") == 4               # 1e-12 <= 1.26e-8 dropped, 1e-6 kept
    assert M.n_keep(s, 1e-10, "rsum2", max_bond=2) == 2
    assert M.n_keep(np.array([1.0, 1e-9]), 0.5, "rsum2") == 1
    assert M.n_keep(np.array([0.0, 0.0]), 1e-10, "rsum2") == 1   # keep at least one
    assert math.isclose(M.renorm_factor(np.array([3.0, 4.0]), 1, 2), 5.0 / 3.0)
    assert M.renorm_factor(np.array([3.0, 4.0]), 2, 2) == 1.0


def test_appendix_a4_probe_table(tensors):
    x = tensors["2d"]
    o = OracleNDMPS.from_tensor(x)
    assert o.bond_sizes() == [34, 512, 64, 8] and o.number_elements_in_MPS() == 615620
    assert np.abs(o.to_tensor() - x).max() <= 2e-13
    a = copy.deepcopy(o)
    a.compress(0.1)
    assert a.bond_sizes() == [34, 461, 64, 8]
    assert abs(a.compression_ratio_on_disk() - 1.49) < 0.01
    b = copy.deepcopy(o)
    b.compress(0.4)
    assert b.bond_sizes() == [34, 149, 32, 1]
    assert abs(b.compression_ratio_on_disk() - 0.39) < 0.01


def test_3d_probe(tensors):
    x = tensors["3d"]
    o = OracleNDMPS.from_tensor(x)
    assert o.bond_sizes() == [160, 136]
    assert np.allclose(o.to_tensor(), x, atol=1e-10)
    o.compress(0.1)
    assert o.bond_sizes() == [160, 1]
    assert abs(o.compression_ratio_on_disk() - 0.015) < 0.002


@pytest.mark.parametrize("mode", ["Std", "DCT"])
def test_reference_invariants(tensors, mode):
    x = tensors["2d"]
    o = OracleNDMPS.from_tensor(x, norm=False, mode=mode)
    assert np.allclose(o.to_tensor(), x, atol=1e-10)                       # test_ndmps.py:35-38
    before = o.number_elements_in_MPS()
    c = copy.deepcopy(o)
    c.compress(0.1)
    assert c.number_elements_in_MPS() < before                             # :46-50
    o.cores[0][:] *= 10                                                    # :53-66
    o.update_boundary_list()
    o.update_norm()
    assert o.boundary_list[0][0] <= o.cores[0].min() and o.boundary_list[0][1] >= o.cores[0].max()
    assert math.isclose(o.norm_value ** 2, M.overlap(o.cores, o.cores), rel_tol=1e-12)


def test_norm_option(tensors):
    o = OracleNDMPS.from_tensor(tensors["2d"], norm=True)                   # test_ndmps.py:41-44
    assert math.isclose(o.norm_value, 1.0, rel_tol=1e-12)


def test_continuous_compress_prints(tensors, capsys):
    o = OracleNDMPS.from_tensor(tensors["2d"][:64, :80])
    o.continuous_compress(0.05, print_ratio=True)
    assert capsys.readouterr().out.count("Compression ratio at") == 20      # test_ndmps.py:75-79


def test_left_canonical_and_truncation_optimality():
    rng = np.random.default_rng(1)
    dims = [6, 5, 4, 7]
    x = rng.standard_normal(dims)
    cores, svals = M.tt_svd(x, dims, max_bond=3, return_svals=True)
    assert M.bond_sizes(cores) == [3, 3, 3]
    a0 = cores[0]
    assert np.allclose(a0.T @ a0, np.eye(3), atol=1e-12)
    a1 = cores[1].reshape(-1, 3)
    assert np.allclose(a1.T @ a1, np.eye(3), atol=1e-12)
    # renorm=2 keeps the Frobenius norm
    assert math.isclose(np.linalg.norm(M.contract_dense(cores)), np.linalg.norm(x), rel_tol=1e-12)
    # overlap of the state with itself equals the squared norm of the dense contraction
    assert math.isclose(M.overlap(cores, cores), np.sum(M.contract_dense(cores) ** 2), rel_tol=1e-12)


def test_compress_bond_is_two_site_svd():
    rng = np.random.default_rng(5)
    t1 = rng.standard_normal((4, 3, 6))
    t2 = rng.standard_normal((6, 5, 2))
    n1, n2, s = M.compress_bond(t1, t2, first=False, last=False, cutoff=0.3)
    theta = t1.reshape(12, 6) @ t2.reshape(6, 10)
    ref = np.linalg.svd(theta, compute_uv=False)
    keep = int(np.sum(ref > 0.3 * ref[0]))
    assert len(s) == keep and np.allclose(s, ref[:keep])
    u, sv, vh = np.linalg.svd(theta, full_matrices=False)
    best = (u[:, :keep] * sv[:keep]) @ vh[:keep]
    assert np.allclose(n1.reshape(12, keep) @ n2.reshape(keep, 10), best, atol=1e-12)


def test_one_site_mps():
    x = np.arange(7.0)
    o = OracleNDMPS.from_tensor(x)
    assert o.bond_sizes() == [] and np.array_equal(o.to_tensor(), x)
    o.compress(0.5)                                         # no-op on L = 1 (core/ndmps.py:103)
    assert np.array_equal(o.to_tensor(), x)
