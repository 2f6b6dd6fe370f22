"""Product host-side planning (imgcompressionmps.utils.core / filetools) against the
reference's golden vectors - the same checks the reference's tests/utils/test_core.py
makes, through the drop-in import path."""
import numpy as np
import pytest

from imgcompressionmps.utils.core import balance_factors, gen_encoding_map, get_factorlist, hierarchical_block_indexing
from imgcompressionmps.utils.filetools import get_num_bits, scale_back, scale_to_dtype

INT64_MAX = 9223372036854775807


def test_balance_factors_reference_cases():
    assert np.prod(balance_factors([2, 2, 3, 3], 2)) == 36
    assert balance_factors([2, 3], 2) == [2, 3]
    assert balance_factors([1, 1, 1, 1], 2) == [1, 1]
    assert np.prod(balance_factors([2] * 10, 2)) == 1024
    assert balance_factors([], 0) == []
    assert balance_factors([6], 1) == [6]
    for bad in (-1, 0):
        with pytest.raises(ValueError):
            balance_factors([2, 3], bad)
    with pytest.raises(ValueError):
        balance_factors([2], 3)


def test_get_factorlist_reference_vectors():
    f, p = get_factorlist((256, 128))
    assert np.array_equal(f, [[2, 2]] * 6 + [[4, 2]])
    assert np.array_equal(p[0], [INT64_MAX, INT64_MAX]) and np.array_equal(p[1:, 0], [128, 64, 32, 16, 8, 4, 1])
    f, p = get_factorlist((30, 40, 50))
    assert np.array_equal(f, [[2, 5, 2], [3, 4, 5], [5, 2, 5]])
    assert np.array_equal(p, [[INT64_MAX] * 3, [15, 8, 25], [5, 2, 5], [1, 1, 1]])
    for bad in ((), (0, 4), (2.0, 3), ("a", "b")):
        with pytest.raises(ValueError):
            get_factorlist(bad)


def test_against_reference_outputs(golden_encoding):
    g = golden_encoding
    for key in sorted({k.split("/")[0] for k in g.files}):
        shape = tuple(int(s) for s in key.split("x"))
        f, p = get_factorlist(shape)
        assert np.array_equal(f, g[f"{key}/factors"]) and np.array_equal(p, g[f"{key}/prod"]), key
        if f"{key}/map" in g.files:
            q, m = gen_encoding_map(shape)
            assert np.array_equal(q, g[f"{key}/qubit_sizes"]) and np.array_equal(m, g[f"{key}/map"]), key
            digits = hierarchical_block_indexing(np.indices(shape), p)
            assert digits.shape == (f.shape[0], len(shape)) + shape


def test_hierarchical_errors():
    with pytest.raises(ValueError):
        hierarchical_block_indexing(np.indices((4, 4)), np.array([[1, 1]]))
    with pytest.raises(ValueError):
        gen_encoding_map(())


def test_quantise_helpers(golden_quantise):
    g = golden_quantise
    assert [get_num_bits(d) for d in (np.uint8, np.uint16, np.int32, np.float32, np.float64)] == list(g["bits"])
    with pytest.raises(ValueError):
        get_num_bits(np.dtype("U4"))
    for i in range(3):
        a = g[f"in{i}"]
        for dt in (np.uint8, np.uint16):
            name = np.dtype(dt).name
            q = scale_to_dtype(a, dt)
            assert np.array_equal(q, g[f"q{i}_{name}"])
            assert np.array_equal(scale_back(q, a.min(), a.max(), dt), g[f"back{i}_{name}"])


def test_volume_pipeline_needs_a_gpu():
    """No CPU fallback: the pipeline refuses to start without a device."""
    import torch
    from imgcompressionmps import _native
    from imgcompressionmps.batch import VolumePipeline
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.NativeError):
        VolumePipeline(workers=2)
    with pytest.raises(ValueError):
        VolumePipeline(workers=0)
