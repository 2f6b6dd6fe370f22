#!/usr/bin/env python
"""One leading-64 eigen-solve of a 512 x 512 Gram matrix, three times: the command behind the per-kernel launch
list of the solver (ncu --metrics gpu__time_duration.sum)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from imgcompressionmps import _native, _ops   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator(device="cuda").manual_seed(3)
a = torch.randn(n, 4 * n, dtype=torch.float64, device="cuda", generator=g)
gm = a @ a.T
for _ in range(3):
    ev, vec, tr, health = _ops.eigh_topk(gm, 64)
torch.cuda.synchronize()
print("health", health, "lambda_1", float(ev[0]))
