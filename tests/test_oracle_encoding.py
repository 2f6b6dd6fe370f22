"""Oracle vs the reference's own golden vectors for the encoding path.

The literal vectors below restate the assertions of the reference's
tests/utils/test_core.py (line numbers cited); the .npz fixtures were produced by
executing the reference's utils/core.py (tests/golden/make_golden.py).
"""
import hashlib

import numpy as np
import pytest

from oracle import encoding as E

INT64_MAX = 9223372036854775807


# ---- balance_factors: reference tests/utils/test_core.py:14-63 ---------------------
def test_balance_preserves_product():
    out = E.balance_factors([2, 2, 3, 3], 2)
    assert np.prod(out) == 36 and len(out) == 2


def test_balance_cases():
    assert E.balance_factors([2, 3], 2) == [2, 3]
    assert E.balance_factors([7, 3, 2, 2], 3) == sorted(E.balance_factors([7, 3, 2, 2], 3))
    assert E.balance_factors([1, 1, 1, 1], 2) == [1, 1]
    big = E.balance_factors([2] * 10, 2)
    assert len(big) == 2 and np.prod(big) == 1024
    assert E.balance_factors([], 0) == []
    assert E.balance_factors([6], 1) == [6]
    with pytest.raises(ValueError):
        E.balance_factors([2, 3], -1)
    with pytest.raises(ValueError):
        E.balance_factors([2, 3], 0)


# ---- get_factorlist: test_core.py:75-88 and :111-122 ------------------------------------
def test_factorlist_256_128():
    f, p = E.get_factorlist((256, 128))
    assert np.array_equal(f, [[2, 2]] * 6 + [[4, 2]])
    assert np.array_equal(p, [[INT64_MAX, INT64_MAX], [128, 64], [64, 32], [32, 16], [16, 8], [8, 4], [4, 2], [1, 1]])


def test_factorlist_30_40_50():
    f, p = E.get_factorlist((30, 40, 50))
    assert np.array_equal(f, [[2, 5, 2], [3, 4, 5], [5, 2, 5]])
    assert np.array_equal(p, [[INT64_MAX] * 3, [15, 8, 25], [5, 2, 5], [1, 1, 1]])


def test_factorlist_misc():
    f, p = E.get_factorlist((30, 24))                       # test_core.py:67-73
    assert f.shape[1] == 2 and p.shape == (f.shape[0] + 1, 2)
    assert np.all(np.prod(f, axis=0) == (30, 24))
    f, p = E.get_factorlist((1, 1))                         # :90-93
    assert np.all(f == 1) and np.all(p >= 1)
    f, _ = E.get_factorlist((4, 6))                         # :95-99 snake
    assert np.all(np.diff(f[:, 1])[::-1] <= 0)
    with pytest.raises(ValueError):                         # :107-109
        E.get_factorlist((0, 4))


# ---- gen_encoding_map: test_core.py:151-173 -----------------------------------------------
def test_encoding_map_8_9():
    q, m = E.gen_encoding_map((8, 9))
    assert np.array_equal(q, [6, 12])
    lvl0 = np.repeat(np.array([[0, 0, 0, 1, 1, 1, 2, 2, 2], [3, 3, 3, 4, 4, 4, 5, 5, 5]]), 4, axis=0)
    row = np.array([[0, 1, 2] * 3, [3, 4, 5] * 3, [6, 7, 8] * 3, [9, 10, 11] * 3])
    lvl1 = np.concatenate([row, row], axis=0)
    assert np.array_equal(m[0], lvl0) and np.array_equal(m[1], lvl1)


def test_encoding_map_errors():
    with pytest.raises(ValueError):
        E.gen_encoding_map(())
    with pytest.raises(ValueError):
        E.gen_encoding_map(("a", "b"))
    q, m = E.gen_encoding_map((1, 4))
    assert m.shape == (len(q), 1, 4)


# ---- hierarchical_block_indexing: test_core.py:212-237 ----------------------------------------
def test_hierarchical_4_6():
    _, prods = E.get_factorlist((4, 6))
    out = E.hierarchical_block_indexing(np.indices((4, 6)), prods)
    assert out.shape == (2, 2, 4, 6)
    assert np.array_equal(out[0, 0], np.repeat([[0], [0], [1], [1]], 6, axis=1))
    assert np.array_equal(out[0, 1], np.tile([0, 0, 1, 1, 2, 2], (4, 1)))
    assert np.array_equal(out[1, 0], np.repeat([[0], [1], [0], [1]], 6, axis=1))
    assert np.array_equal(out[1, 1], np.tile([0, 1, 0, 1, 0, 1], (4, 1)))
    with pytest.raises(ValueError):
        E.hierarchical_block_indexing(np.indices((4, 4)), np.array([[1, 1]]))


# ---- fixtures produced by running the reference -------------------------------------------------
def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _shapes(golden):
    return sorted({k.split("/")[0] for k in golden.files})


def test_against_reference_outputs(golden_encoding):
    g = golden_encoding
    for key in _shapes(g):
        shape = tuple(int(s) for s in key.split("x"))
        f, p = E.get_factorlist(shape)
        assert np.array_equal(f, g[f"{key}/factors"]), key
        assert np.array_equal(p, g[f"{key}/prod"]), key
        if f"{key}/qubit_sizes" not in g.files:
            continue
        assert E.site_dims(shape) == list(g[f"{key}/qubit_sizes"]), key
        ramp = np.arange(int(np.prod(shape)), dtype=np.int32).reshape(shape)
        dense = E.encode(ramp)
        if f"{key}/encoded_ramp" in g.files:
            assert np.array_equal(dense, g[f"{key}/encoded_ramp"]), key
            q, m = E.gen_encoding_map(shape)
            assert np.array_equal(m, g[f"{key}/map"]), key
        else:
            assert _sha(dense) == str(g[f"{key}/encoded_ramp_sha256"]), key
        assert np.array_equal(E.decode(dense, shape), ramp), key


def test_encode_equals_scatter_through_map():
    rng = np.random.default_rng(3)
    for shape in [(8, 9), (12, 18, 10), (4, 4, 4, 4), (6, 10, 15), (7,)]:
        x = rng.random(shape)
        assert np.array_equal(E.encode(x), E.encode_by_map(x))
