#!/usr/bin/env python
"""Stage times of one 256^3 phantom volume for every chi of BASELINE configs[1] (128 ... 8)."""
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

ctx = _native.context()
vol = torch.from_numpy(synthetic_volume((256, 256, 256), 2026)).cuda()
for chi in (128, 64, 32, 16, 8):
    obj = NDMPS.from_tensor(vol, max_bond=chi)
    obj.to_tensor_device()
    ctx.profile(True)
    ctx.stage_times(reset=True)
    ctx.stat("launches", reset=True) if False else None
    for _ in range(3):
        o2 = NDMPS.from_tensor(vol, max_bond=chi)
        o2.to_tensor_device()
    torch.cuda.synchronize()
    st = ctx.stage_times(reset=True)
    ctx.profile(False)
    print(f"chi={chi}: bonds {obj.bond_sizes()}  " + " ".join(f"{k}={v[0] / 3:.2f}ms/{int(v[1] / 3)}" for k, v in st.items() if v[1]), flush=True)
