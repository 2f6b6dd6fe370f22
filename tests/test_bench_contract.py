"""The JSON line of bench.py's reference arm (the CPU leg runs anywhere) carries the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "voxels/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # both arms print the SAME config object: it is a function of the command line only
    sys.path.insert(0, str(ROOT))
    import argparse

    import bench
    args = argparse.Namespace(chi=0)
    assert d["config"] == bench.workload_config("cfg1", args, 1)
    assert bench.workload_config("cfg3", args, 1)["shape"] == [512, 512, 512]


def test_default_workload_is_the_north_star_volume_and_cfg5_is_reported_degenerate():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", "cfg5"], capture_output=True, text=True, timeout=300,
                         cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][0])
    assert d["degenerate"]["levels"] == 1 and d["degenerate"]["site_dims"] == [1920 * 1080 * 3 * 512] and d["degenerate"]["bonds"] == []
    sys.path.insert(0, str(ROOT))
    import bench
    src = (ROOT / "bench.py").read_text()
    assert 'ap.add_argument("--workload", default="cfg3"' in src
    assert bench.WORKLOADS["cfg3"]["shape"] == (512, 512, 512) and bench.WORKLOADS["cfg3"]["chi"] == 64


def test_algorithmic_work_matches_survey_table():
    """256^3 at chi = 64 -> 70.9 bytes/voxel and 1895 flop/voxel for encode + sweep + reconstruct + decode."""
    sys.path.insert(0, str(ROOT))
    import bench
    dims, ranks = [8] * 8, [8, 64, 64, 64, 64, 64, 8]
    w = bench.algorithmic_work(dims, ranks)
    n = 8 ** 8
    bytes_per_voxel = (w["encode_bytes"] + w["decode_bytes"] + w["sweep_bytes"] + w["recon_bytes"]) / n
    flops_per_voxel = (w["gram_flops"] + w["project_flops"] + w["recon_flops"]) / n
    assert abs(bytes_per_voxel - 70.9) < 0.2
    assert abs(flops_per_voxel - 1895) < 5
    assert w["gram_flops_issued"] < w["gram_flops_executed"]


def test_roofline_entries_follow_the_contract():
    """Every stage entry: ``bound`` from the contract's enum, ``frac`` = achieved / peak, ``traffic`` read from an ncu CSV
    committed under profiles/ (the row of the kernel that really runs the stage), algorithmic figures beside it."""
    import bench
    info = {"site_dims": [8] * 9, "bond_dims": [8, 37, 64, 64, 64, 64, 64, 8]}
    nvox = 512 ** 3
    work = bench.work_for("cfg3", info, nvox)
    per_call = {"permute": (0.36, 2), "gram": (1.0, 6), "eig": (8.2, 8), "project": (0.32, 4), "contract": (0.23, 1)}
    peaks = {"hbm_gbs": 6559.4, "bf16_tflops": 1648.0}
    roof = bench.build_rooflines(per_call, 11.2, work, nvox, peaks, 35.3, 1.1e9, True, True)
    assert set(roof) == set(per_call)
    for stage, e in roof.items():
        assert e["bound"] in ("hbm", "tensor"), stage
        assert e["frac"] == pytest.approx(e["achieved"] / e["peak"])
        assert e["traffic"] and e["traffic_source"].startswith("profiles/r0"), stage
        assert e["share_of_tensor_time"] == pytest.approx(per_call[stage][0] / 11.2)
    # the permutation of a power-of-two volume is the register kernel: 2 x 8 B/voxel in 0.36 ms against the measured copy peak
    assert "permute_bits_kernel" in roof["permute"]["kernel"] and roof["permute"]["traffic_source"].endswith("r02b_kernels_raw.csv")
    assert roof["permute"]["algorithmic_bytes_per_tensor"] == 16 * nvox and roof["permute"]["frac"] == pytest.approx(0.909, abs=2e-3)
    # ncu's DRAM traffic of ONE launch (encode or decode: 8 B/voxel) = the algorithmic bytes of that launch
    assert roof["permute"]["calls_per_tensor"] == 2 and roof["permute"]["traffic"] == pytest.approx(8 * nvox, rel=0.06)
    assert "proj_i8_kernel" in roof["project"]["kernel"] and roof["project"]["traffic"] > 6e8
    assert "latency-bound" in roof["eig"]["bound_note"]
    tiles = bench.build_rooflines(per_call, 11.2, work, nvox, peaks, 35.3, 1.1e9, True, False)
    assert "permute_tiled_kernel" in tiles["permute"]["kernel"]
