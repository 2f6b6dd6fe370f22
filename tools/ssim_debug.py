#!/usr/bin/env python
"""float32 SSIM arithmetic vs float64 per slice on the 64^3 phantom of tests/test_gpu_ndmps.py: where is the worst slice?"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
from conftest import phantom
from imgcompressionmps import _native as N, _ops
from imgcompressionmps.core.ndmps import NDMPS
ctx = N.context()
x = phantom((64, 64, 64), seed=7).astype(np.float32)
rec = NDMPS.from_tensor(x, max_bond=16).to_tensor()
xd, rd = torch.from_numpy(x).cuda(), torch.from_numpy(rec).cuda()
for ax in range(3):
    ctx.set_option("ssim_exact", 1); e = _ops.ssim_slices(xd, rd, ax)
    ctx.set_option("ssim_exact", 0); f = _ops.ssim_slices(xd, rd, ax)
    d = np.abs(e - f)
    i = int(np.nanargmax(d))
    sl = [slice(None)] * 3; sl[ax] = i
    xs, rs = x[tuple(sl)], np.clip(rec[tuple(sl)], 0, None)
    print(f"axis {ax}: worst slice {i}: exact {e[i]:.9f} fast {f[i]:.9f} diff {d[i]:.2e}; orig range [{xs.min():.3e}, {xs.max():.3e}] rec range [{rs.min():.3e}, {rs.max():.3e}]; "
          f"slices over 5e-6: {int((d > 5e-6).sum())}, over 1e-6: {int((d > 1e-6).sum())}")
