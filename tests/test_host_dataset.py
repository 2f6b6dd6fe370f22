"""Host-side rows of SURVEY section 8(f): file helpers, the NIfTI / .npz loader, the curve aggregation.
Everything here runs without a GPU; the reference's own helpers (importable: numpy only) generated the
expectations where the reference can run (`filetools.py`), the rest is checked against its definition."""
import gzip
import json
import struct
import sys
from pathlib import Path

import numpy as np
import pytest

from imgcompressionmps.utils import filetools as FT
from imgcompressionmps.utils.metrics import compute_mean_std


def test_mri_to_slices_and_shapes(capsys):
    rng = np.random.default_rng(0)
    vols = [rng.random((6, 8, 10)), rng.random((4, 4)), rng.random((5, 7, 9))]
    slices, bits = FT.mri_to_slices(vols, [16, 8, 32])
    assert "Skipping non-3D volume at index 1" in capsys.readouterr().out
    assert [s.shape for s in slices] == [(8, 10), (6, 10), (6, 8), (7, 9), (5, 9), (5, 7)]
    assert np.array_equal(slices[0], vols[0][3]) and np.array_equal(slices[1], vols[0][:, 4]) and np.array_equal(slices[5], vols[2][:, :, 4])
    assert bits == [16, 16, 16, 32, 32, 32]
    assert FT.mri_to_slices(vols)[1] == [16] * 6
    assert FT.get_shapes(vols) == [(6, 8, 10), (4, 4), (5, 7, 9)]


def test_find_files_and_combine_jsons(tmp_path, capsys):
    (tmp_path / "a").mkdir()
    for name in ("a/x.npz", "a/y.gz", "z.npz"):
        (tmp_path / name).write_bytes(b"")
    assert sorted(Path(p).name for p in FT.find_specific_files(tmp_path, ".npz")) == ["x.npz", "z.npz"]
    assert len(FT.find_specific_files(tmp_path)) == 3
    one = {"mode": "DCT", "cutoff_list": [0.1, 0.2], "ssim": [[1, 2]], "files": ["f1"], "datatype": "MRI"}
    two = {"mode": "DCT", "cutoff_list": [0.1, 0.2], "ssim": [[3, 4]], "files": ["f2"], "datatype": "MRI"}
    (tmp_path / "1.json").write_text(json.dumps(one))
    (tmp_path / "2.json").write_text(json.dumps(two))
    FT.combine_jsons(tmp_path / "1.json", tmp_path / "2.json", tmp_path / "out.json")
    merged = json.loads((tmp_path / "out.json").read_text())
    assert merged == {"mode": "DCT", "cutoff_list": [0.1, 0.2], "ssim": [[1, 2], [3, 4]], "files": ["f1", "f2"], "datatype": "MRI"}
    assert "Combined JSON created successfully!" in capsys.readouterr().out
    assert (FT.find_project_root("imgcompressionmps") / "imgcompressionmps").exists()
    with pytest.raises(FileNotFoundError):
        FT.find_project_root("no-such-marker-directory")


def _write_nifti(path, data, slope=1.0, inter=0.0):
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    dims = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into("<8h", hdr, 40, *dims)
    code = {np.dtype(np.int16): 4, np.dtype(np.float32): 16, np.dtype(np.uint8): 2}[data.dtype]
    struct.pack_into("<h", hdr, 70, code)
    struct.pack_into("<h", hdr, 72, data.dtype.itemsize * 8)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, slope, inter)
    hdr[344:348] = b"n+1\0"
    with gzip.open(path, "wb") as f:
        f.write(bytes(hdr) + data.tobytes(order="F"))


def test_load_tensors_npz_and_nifti(tmp_path, capsys):
    from imgcompressionmps.evaluation.loader import load_tensors, read_nifti
    rng = np.random.default_rng(1)
    seq = rng.random((6, 10, 12)).astype(np.float32)
    np.savez(tmp_path / "clip.npz", sequence=seq)
    data, bits = load_tensors([str(tmp_path / "clip.npz")], ".npz")
    assert "Loading file 1/1" in capsys.readouterr().out
    assert bits == [32] and np.array_equal(data[0], seq)
    cropped, _ = load_tensors([str(tmp_path / "clip.npz")], ".npz", shape=(4, 5, 6))
    assert cropped[0].shape == (4, 5, 6) and np.array_equal(cropped[0], seq[:4, :5, :6])
    vol = rng.integers(-500, 3000, size=(5, 6, 7)).astype(np.int16)
    _write_nifti(tmp_path / "scan.nii.gz", vol, slope=2.0, inter=-1.0)
    got, stored = read_nifti(tmp_path / "scan.nii.gz")
    assert stored == np.dtype(np.int16) and got.dtype == np.float64 and np.array_equal(got, vol.astype(np.float64) * 2.0 - 1.0)
    data, bits = load_tensors([str(tmp_path / "scan.nii.gz")], ".gz")
    assert bits == [16] and data[0].shape == (5, 6, 7)
    with pytest.raises(ValueError):
        load_tensors([], ".nii")


def test_compute_mean_std_matches_the_definition():
    rng = np.random.default_rng(2)
    n_samples, n_levels = 4, 6
    ratios = np.sort(rng.uniform(0.05, 0.9, (n_samples, n_levels)), axis=1)[:, ::-1].copy()   # ratio falls as the cutoff grows
    ssim = np.sort(rng.uniform(0.5, 1.0, (n_samples, n_levels)), axis=1)[:, ::-1].copy()
    shapes = [(64, 64, 32), (7, 11, 13), (30, 40, 50), (128, 128, 3)]          # sample 1 is all primes: excluded
    d = {"compressionratio_list_disk": ratios.tolist(), "ssim_list": ssim.tolist(), "shapes": shapes}
    mean, std, grid = compute_mean_std(d, 9)
    lo, hi = np.max(1 / ratios[:, 0]), np.min(1 / ratios[:, -1])
    assert np.allclose(grid, np.linspace(lo, hi, 9))
    want = [np.interp(grid, 1 / ratios[i], ssim[i]) for i in (0, 2, 3)]
    assert np.allclose(mean, np.mean(want, axis=0)) and np.allclose(std, np.std(want, axis=0))
    only_primes = {"x": ratios[:1].tolist(), "y": ssim[:1].tolist(), "shapes": [(7, 11, 13)]}
    m, s, g = compute_mean_std(only_primes, 5, key_x="x", key_y="y")
    assert np.isnan(m) and np.isnan(s) and len(g) == 5


def test_benchmark_module_exports_the_reference_names():
    import imgcompressionmps.evaluation.benchmark as B
    for name in ("load_tensors", "conv_to_mps", "conv_to_tensors", "compress_list", "benchmark_metric", "run_benchmark",
                 "run_full_benchmark"):
        assert callable(getattr(B, name)), name
    with pytest.raises(FileNotFoundError):
        B.run_full_benchmark(Path(__file__).parent / "golden", np.array([0.1]), "r.json", ending=".does-not-exist")
