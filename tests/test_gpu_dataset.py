"""SURVEY section 8(f) rows on the device: data-set loop (`run_full_benchmark`, prefetching uploads) and
quantisation to every integer dtype `np.iinfo` knows (`core/ndmps.py:182-207`, `utils/filetools.py:20-39`)."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from conftest import phantom                              # noqa: E402
from oracle import quantise as OQ                         # noqa: E402
from oracle.ndmps import OracleNDMPS                      # noqa: E402


def test_run_full_benchmark_writes_the_reference_layout(tmp_path, monkeypatch, capsys):
    from imgcompressionmps.evaluation.benchmark import run_full_benchmark
    vols = [phantom((16, 24, 20), seed=80 + i, background=0.01) for i in range(3)]
    for i, v in enumerate(vols):
        np.savez(tmp_path / f"clip{i}.npz", sequence=v)
    monkeypatch.chdir(tmp_path)
    cutoffs = np.array([0.05, 0.2])
    run_full_benchmark(tmp_path, cutoffs, "run.json", datatype="Video", mode="Std", ending=".npz")
    out = capsys.readouterr().out
    assert "Loading file 3/3" in out and "Converting file 3/3" in out and "Starting benchmark..." in out
    res = json.loads((tmp_path / "src/evaluation/results/run.json").read_text())
    assert set(res) == {"datatype", "mode", "files", "cutoff_list", "bitsize_list", "shapes", "ssim", "compression_ratio",
                        "bond_dims", "psnr", "fidelity", "storage", "gzip_bytes", "gzip_ratio"}
    assert res["datatype"] == "Video" and res["mode"] == "Std" and res["cutoff_list"] == [0.05, 0.2]
    assert res["bitsize_list"] == [64, 64, 64] and res["shapes"] == [[16, 24, 20]] * 3 and len(res["files"]) == 3
    assert np.asarray(res["ssim"]).shape == (3, 3) and len(res["bond_dims"]) == 3
    # compression ratios depend on bond dimensions only: exact against the oracle's loop, file by file
    order = [int(f[-5]) for f in res["files"]]
    for row, idx in zip(res["compression_ratio"], order):
        o = OracleNDMPS.from_tensor(vols[idx])
        want = [o.compression_ratio()]
        for c in cutoffs:
            o.compression_ratio_on_disk(np.uint16, replace=True)      # the gzip_ratio metric quantises in place before each cut
            o.compress(float(c))
            want.append(o.compression_ratio())
        assert row[0] == want[0]
        assert np.allclose(row, want, rtol=0.25)          # later levels see gauge-dependent quantisation noise
        assert all(b <= a for a, b in zip(row, row[1:]))
    # MRI_Slice: three central slices per volume
    run_full_benchmark(tmp_path, cutoffs, "slices.json", datatype="MRI_Slice", mode="DCT", ending=".npz", start=0, end=1)
    res = json.loads((tmp_path / "src/evaluation/results/slices.json").read_text())
    assert res["shapes"] == [[24, 20], [16, 20], [16, 24]] and len(res["files"]) == 1


def test_prefetch_to_device_order_and_dtypes():
    from imgcompressionmps.evaluation.loader import prefetch_to_device
    rng = np.random.default_rng(3)
    arrays = [rng.random((40, 50)).astype(np.float32), rng.random((40, 50)), rng.integers(0, 9, (7, 9, 11)).astype(np.int16),
              rng.random((40, 50)).astype(np.float32)]
    got = list(prefetch_to_device(iter(arrays), depth=2))
    assert [g.dtype for g in got] == [torch.float32, torch.float64, torch.float64, torch.float32]
    for g, a in zip(got, arrays):
        assert g.is_cuda and np.array_equal(g.cpu().numpy(), a.astype(g.cpu().numpy().dtype))


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32])
def test_compress_to_dtype_any_integer_dtype(dtype):
    """Bytes equal numpy's scale_to_dtype on the same core values (host copy of the device cores)."""
    from imgcompressionmps.core.ndmps import NDMPS
    x = phantom((16, 16, 16), seed=9)
    g = NDMPS.from_tensor(x, max_bond=6)
    cores = [np.array(c) for c in g.mps.arrays]
    ints = g.compress_to_dtype(dtype)
    for q, c in zip(ints, cores):
        assert q.dtype == np.dtype(dtype) and q.shape == c.shape
        assert np.array_equal(q, OQ.scale_to_dtype(c, dtype))
    before = g.to_tensor()
    g.compress_to_dtype(dtype, replace=True)
    want = [OQ.scale_back(OQ.scale_to_dtype(c, dtype), c.min(), c.max(), dtype) for c in cores]
    for got, w in zip(g.mps.arrays, want):
        assert np.allclose(np.asarray(got), w, rtol=0, atol=1e-12 * max(1.0, np.abs(w).max()))
    tol = {8: 0.2, 16: 1e-3}.get(np.iinfo(dtype).bits, 1e-6)
    assert np.abs(g.to_tensor() - before).max() < tol
