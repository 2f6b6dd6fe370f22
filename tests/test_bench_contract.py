"""The JSON line of bench.py's reference arm (the CPU leg runs anywhere) carries the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

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
