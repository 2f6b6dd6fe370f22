#!/usr/bin/env python
"""One 512^3 (or N^3) float32 volume through NDMPS.from_tensor(max_bond=64) + to_tensor, twice: the command behind
the per-volume launch list (ncu --metrics gpu__time_duration.sum)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from bench import synthetic_volume                     # noqa: E402
from imgcompressionmps import _native                  # noqa: E402
from imgcompressionmps.core.ndmps import NDMPS         # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = torch.from_numpy(synthetic_volume((n, n, n), 2027)).cuda()
ctx = _native.context()
for it in range(2):
    ctx.stat("launches", reset=True) if False else None
    obj = NDMPS.from_tensor(vol, max_bond=64)
    rec = obj.to_tensor_device()
torch.cuda.synchronize()
print("bonds", obj.bond_sizes(), "rel err", float(torch.linalg.vector_norm((rec - vol).double()) / torch.linalg.vector_norm(vol.double())))
