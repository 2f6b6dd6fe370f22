"""Run a few eigensolves of one size (for ncu): python tools/eig_probe.py <n> [reps]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "img-compression-mps_b200"))
from imgcompressionmps import _ops  # noqa: E402
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(n, 4 * n, dtype=torch.float64, device="cuda", generator=g)
gm = a @ a.T
for _ in range(reps):
    ev, vec, sw = _ops.eigh(gm)
torch.cuda.synchronize()
print("n", n, "sweeps", sw, "ok", bool(torch.isfinite(ev).all()))
