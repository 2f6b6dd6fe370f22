"""Oracle vs outputs of the REFERENCE'S OWN ``core/ndmps.py``, ``utils/metrics.py`` and ``evaluation/benchmark.py``,
executed in the build container with stand-ins for quimb / scikit-image built on the oracle's restatement of those
libraries (``tests/golden/make_golden_reference_exec.py`` -> ``reference_class.npz``, ``reference_metrics.npz``,
``reference_benchmark.npz``).

Pinned by these fixtures: everything the reference's code does around the third-party cores - scatter / gather through
its encoding map, norm / DCT options, boundary list, norm value, the compress loop, counts and ratios, quantisation with
the stored boundaries, gzip sizes, storage, the lines ``continuous_compress`` prints; metric clipping, data range, window
choice, slice order, axis / frame averaging, dispatch, PSNR, ``compute_mean_std``.  NOT pinned: the inside of the TT-SVD,
the bond compression, the overlap and the SSIM window formula (the oracle stands on both sides there)."""
import contextlib
import io

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import metrics as OM
from oracle.ndmps import OracleNDMPS

CASES = ["rand2d", "rand3d_dct", "rand3d_norm", "smooth3d", "smooth4d_dct_norm"]
TIGHT = dict(rtol=1e-12, atol=1e-13)


@pytest.fixture(scope="module")
def ref_class():
    return np.load(GOLDEN / "reference_class.npz")


@pytest.fixture(scope="module")
def ref_metrics():
    return np.load(GOLDEN / "reference_metrics.npz")


def flat(arrs):
    return np.concatenate([np.ravel(a) for a in arrs])


@pytest.mark.parametrize("name", CASES)
def test_oracle_class_equals_the_executed_reference_class(ref_class, name):
    g = ref_class
    x = g[f"{name}/input"]
    norm, mode, cutoff = (str(v) for v in g[f"{name}/options"])
    norm, cutoff = norm == "True", float(cutoff)
    o = OracleNDMPS.from_tensor(x.copy(), norm=norm, mode=mode)
    # core/ndmps.py:36-78, 88-92, 127-129, 159-161
    assert list(o.qubit_size) == list(g[f"{name}/qubit_size"])
    assert o.bond_sizes() == list(g[f"{name}/bonds0"])
    assert np.allclose(o.boundary_list, g[f"{name}/boundary0"], **TIGHT)
    assert o.norm_value == pytest.approx(float(g[f"{name}/norm0"]), rel=1e-13)
    assert o.compression_ratio() == float(g[f"{name}/ratio0"])
    assert o.number_elements_in_MPS() == int(g[f"{name}/elements0"])
    assert o.get_storage_space(np.uint16) == float(g[f"{name}/storage0"])
    # core/ndmps.py:131-153: contraction, gather through the map, inverse DCT; lossless up to the 1e-10 trim
    assert np.allclose(o.to_tensor(), g[f"{name}/tensor0"], **TIGHT)
    scale = np.linalg.norm(x) if norm else 1.0
    assert np.allclose(o.to_tensor() * scale, x, atol=1e-8)
    # core/ndmps.py:94-108
    o.compress(cutoff)
    assert o.bond_sizes() == list(g[f"{name}/bonds1"])
    assert np.allclose(o.boundary_list, g[f"{name}/boundary1"], **TIGHT)
    assert o.norm_value == pytest.approx(float(g[f"{name}/norm1"]), rel=1e-13)
    assert o.compression_ratio() == float(g[f"{name}/ratio1"])
    assert np.allclose(o.to_tensor(), g[f"{name}/tensor1"], **TIGHT)
    # core/ndmps.py:182-277 + utils/filetools.py:7-39: integer payloads and byte counts are exact
    for dt in (np.uint16, np.uint8):
        tag = np.dtype(dt).name
        ints = o.compress_to_dtype(dt)
        assert all(a.dtype == dt for a in ints)
        assert np.array_equal(flat(ints), g[f"{name}/ints_{tag}"])
        assert o.get_bytesize_on_disk(dt) == int(g[f"{name}/gzip_{tag}"])
        assert o.compression_ratio_on_disk(dt) == float(g[f"{name}/disk_ratio_{tag}"])
        assert o.get_storage_space(dt) == float(g[f"{name}/storage_{tag}"])
    # quantise in place with the STORED boundaries (core/ndmps.py:201-206), refreshed boundary list and norm
    o.compress_to_dtype(np.uint8, replace=True)
    assert np.allclose(o.boundary_list, g[f"{name}/boundary2"], **TIGHT)
    assert o.norm_value == pytest.approx(float(g[f"{name}/norm2"]), rel=1e-13)
    assert np.allclose(o.to_tensor(), g[f"{name}/tensor2"], **TIGHT)
    # core/ndmps.py:110-125: twenty compressions, twenty printed lines with the same ratios
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        o.continuous_compress(2 * cutoff)
    assert buf.getvalue() == str(g[f"{name}/printed"])
    assert len(buf.getvalue().splitlines()) == 20
    assert o.bond_sizes() == list(g[f"{name}/bonds3"])
    assert np.allclose(o.to_tensor(), g[f"{name}/tensor3"], rtol=1e-10, atol=1e-12)
    # utils/metrics.py:149-160
    fresh = OracleNDMPS.from_tensor(x.copy(), norm=norm, mode=mode)
    fid = OM.compute_overlap(o.cores, o.norm_value, fresh.cores, fresh.norm_value)
    assert fid == pytest.approx(float(g[f"{name}/fidelity"]), rel=1e-12)


@pytest.mark.parametrize("name", ["img", "small", "tiny", "vol", "thin", "series"])
def test_oracle_metrics_equal_the_executed_reference_metrics(ref_metrics, name):
    m = ref_metrics
    a, b = m[f"{name}/a"], m[f"{name}/b"]
    assert b.min() < 0                                              # the clip at utils/metrics.py:23 is exercised
    assert OM.compute_ssim_by_dim(a, b) == pytest.approx(float(m[f"{name}/ssim"]), rel=1e-13)
    assert OM.compute_ssim_by_dim(b, a) == pytest.approx(float(m[f"{name}/ssim_swapped"]), rel=1e-13)
    assert OM.compute_psnr(a, b) == pytest.approx(float(m[f"{name}/psnr"]), rel=1e-13)
    assert OM.compute_psnr(a, a) == np.inf and np.isinf(m[f"{name}/psnr_same"])
    if a.ndim == 3:
        for ax in range(3):
            assert np.allclose(OM.ssim_3d_axis(a, b, ax), m[f"{name}/axis{ax}"], rtol=1e-13, atol=0)
        assert OM.ssim_3d_axis(a, b, -1) == [] and m[f"{name}/axis_neg"].size == 0     # the reference's quirk


def test_package_compute_mean_std_equals_the_executed_reference(ref_metrics):
    """utils/metrics.py:163-202 is host-only post-processing: the package's own function, no GPU."""
    from imgcompressionmps.utils.metrics import compute_mean_std
    m = ref_metrics
    curves = {"compressionratio_list_disk": m["mean_std/x"].tolist(), "ssim_list": m["mean_std/y"].tolist(),
              "shapes": m["mean_std/shapes"].tolist()}
    mean, std, grid = compute_mean_std(curves, 6)
    assert np.allclose(mean, m["mean_std/mean"], **TIGHT)
    assert np.allclose(std, m["mean_std/std"], **TIGHT)
    assert np.allclose(grid, m["mean_std/grid"], **TIGHT)
    primes = {"compressionratio_list_disk": curves["compressionratio_list_disk"][:1], "ssim_list": curves["ssim_list"][:1],
              "shapes": [[7, 11, 13]]}
    m2, s2, g2 = compute_mean_std(primes, 4)
    assert np.isnan(m2) and np.isnan(s2) and bool(m["mean_std/all_prime_is_nan"][0])
    assert np.allclose(g2, m["mean_std/all_prime_grid"], **TIGHT)


def test_oracle_cutoff_sweep_equals_the_executed_reference_run_benchmark():
    """evaluation/benchmark.py:149-194 executed from the reference's source (stand-ins as above) on three volumes and two
    cutoffs, against the same loop written on the oracle - the loop the GPU test of the package's ``run_benchmark``
    compares with (tests/test_gpu_ndmps.py).  Pins the order of the metric calls inside a level, in particular that the
    ``gzip_ratio`` metric quantises the cores IN PLACE (uint16) after the level's SSIM / PSNR / fidelity were taken, and
    that the SECOND argument of the SSIM call - the original - is the one clipped at 0."""
    import copy
    b = np.load(GOLDEN / "reference_benchmark.npz")
    vols = [v for v in b["volumes"]]
    assert min(v.min() for v in vols) < 0                       # originals with negative voxels: the clip matters
    cutoffs = [float(c) for c in b["cutoffs"]]
    objs = [OracleNDMPS.from_tensor(v.copy(), norm=False, mode="Std") for v in vols]
    refs = copy.deepcopy(objs)
    want = {k: [] for k in ("ssim", "compression_ratio", "bond_dims", "psnr", "fidelity", "storage", "gzip_bytes", "gzip_ratio")}
    for level in [None] + cutoffs:
        if level is not None:
            for o in objs:
                o.compress(level)
        want["ssim"].append([OM.compute_ssim_by_dim(o.to_tensor(), v) for o, v in zip(objs, vols)])
        want["compression_ratio"].append([o.compression_ratio() for o in objs])
        want["bond_dims"].append([o.bond_sizes() for o in objs])
        want["psnr"].append([OM.compute_psnr(o.to_tensor(), v) for o, v in zip(objs, vols)])
        want["fidelity"].append([OM.compute_overlap(o.cores, o.norm_value, r.cores, r.norm_value) for o, r in zip(objs, refs)])
        want["storage"].append([o.get_storage_space(np.uint16) for o in objs])
        want["gzip_bytes"].append([o.get_bytesize_on_disk(dtype=np.uint16) for o in objs])
        want["gzip_ratio"].append([o.compression_ratio_on_disk(dtype=np.uint16, replace=True) for o in objs])
    assert np.array_equal(np.array(want["bond_dims"]), b["bond_dims"])
    for key in ("compression_ratio", "storage", "gzip_bytes", "gzip_ratio"):
        assert np.array_equal(np.array(want[key]).T, b[key]), key
    for key in ("ssim", "psnr", "fidelity"):
        assert np.allclose(np.array(want[key]).T, b[key], rtol=1e-11, atol=1e-12), key
    assert b["ssim"].shape == (3, 3) and np.all(b["ssim"][:, 0] < 1.0)      # lossless level, SSIM below 1: the clipped original
    assert str(b["printed"]).count("Converting file") == 3 and "Status: 100.00% - Cutoff: 0.1" in str(b["printed"])
