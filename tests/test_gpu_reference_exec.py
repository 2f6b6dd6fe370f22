"""The CUDA path against outputs of the REFERENCE'S OWN ``core/ndmps.py`` / ``utils/metrics.py`` executed in the build
container (``tests/golden/make_golden_reference_exec.py``; quimb / scikit-image replaced there by stand-ins built on the
oracle's restatement).  float64 inputs, the reference's defaults.  Gauge-dependent quantities (core entries, quantised
payloads, gzip sizes) are compared where the reference's tests do: as ratios with a tolerance."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CASES = ["rand2d", "rand3d_dct", "rand3d_norm", "smooth3d", "smooth4d_dct_norm"]


@pytest.fixture(scope="module")
def ref_class():
    return np.load(GOLDEN / "reference_class.npz")


@pytest.fixture(scope="module")
def ref_metrics():
    return np.load(GOLDEN / "reference_metrics.npz")


@pytest.mark.parametrize("name", CASES)
def test_ndmps_class_against_the_executed_reference(ref_class, name):
    from imgcompressionmps.core.ndmps import NDMPS
    from imgcompressionmps.utils.metrics import compute_overlap
    g = ref_class
    x = g[f"{name}/input"]
    norm, mode, cutoff = (str(v) for v in g[f"{name}/options"])
    norm, cutoff = norm == "True", float(cutoff)
    obj = NDMPS.from_tensor(x.copy(), norm=norm, mode=mode)
    assert [int(q) for q in obj.qubit_size] == list(g[f"{name}/qubit_size"])
    assert obj.bond_sizes() == list(g[f"{name}/bonds0"])
    assert obj.number_elements_in_MPS() == int(g[f"{name}/elements0"])
    assert obj.compression_ratio() == float(g[f"{name}/ratio0"])
    assert obj.get_storage_space(np.uint16) == float(g[f"{name}/storage0"])
    assert float(obj.norm_value) == pytest.approx(float(g[f"{name}/norm0"]), rel=1e-10)
    assert np.allclose(obj.to_tensor(), g[f"{name}/tensor0"], rtol=0, atol=1e-9)
    obj.compress(cutoff)
    assert obj.bond_sizes() == list(g[f"{name}/bonds1"])
    assert obj.compression_ratio() == float(g[f"{name}/ratio1"])
    assert float(obj.norm_value) == pytest.approx(float(g[f"{name}/norm1"]), rel=1e-8)
    assert np.allclose(obj.to_tensor(), g[f"{name}/tensor1"], rtol=0, atol=1e-7)
    fresh = NDMPS.from_tensor(x.copy(), norm=norm, mode=mode)
    assert float(compute_overlap(obj, fresh)) == pytest.approx(float(g[f"{name}/fidelity1"]), rel=1e-8)
    for dt in (np.uint16, np.uint8):
        tag = np.dtype(dt).name
        assert obj.get_storage_space(dt) == float(g[f"{name}/storage_{tag}"])
        ints = obj.compress_to_dtype(dt)
        assert all(a.dtype == dt for a in ints) and sum(a.size for a in ints) == g[f"{name}/ints_{tag}"].size
        # gzip of gauge-dependent integers: same size class, not the same bytes
        assert obj.get_bytesize_on_disk(dt) == pytest.approx(int(g[f"{name}/gzip_{tag}"]), rel=0.25)


@pytest.mark.parametrize("name", ["img", "small", "tiny", "vol", "thin", "series"])
def test_metrics_against_the_executed_reference(ref_metrics, name):
    from imgcompressionmps.utils.metrics import compute_psnr, compute_ssim_by_dim, ssim_3d_axis
    m = ref_metrics
    a, b = m[f"{name}/a"], m[f"{name}/b"]
    assert compute_ssim_by_dim(a, b) == pytest.approx(float(m[f"{name}/ssim"]), abs=1e-9)
    assert compute_ssim_by_dim(b, a) == pytest.approx(float(m[f"{name}/ssim_swapped"]), abs=1e-9)
    assert compute_psnr(a, b) == pytest.approx(float(m[f"{name}/psnr"]), abs=1e-9)
    assert compute_psnr(a, a) == np.inf
    if a.ndim == 3:
        for ax in range(3):
            assert np.allclose(ssim_3d_axis(a, b, ax), m[f"{name}/axis{ax}"], rtol=0, atol=1e-9)
        assert ssim_3d_axis(a, b, -1) == []
