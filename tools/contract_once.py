#!/usr/bin/env python
"""The final product of a 512^3 reconstruction alone (dense = X W, 32768 x 64 times 64 x 4096, float32): the command
behind the ncu capture of gemm_tc_kernel."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from imgcompressionmps import _native, _ops   # noqa: E402

g = torch.Generator(device="cuda").manual_seed(5)
x = torch.rand((32768, 64), device="cuda", generator=g, dtype=torch.float32) - 0.3
w = torch.rand((64, 4096), device="cuda", generator=g, dtype=torch.float32) - 0.5
ctx = _native.context()
ctx.set_option("gemm_path", 3)
for _ in range(3):
    c = _ops.gemm(x, w, out_dtype=torch.float32)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    c = _ops.gemm(x, w, out_dtype=torch.float32)
e.record()
torch.cuda.synchronize()
err = float((c.double() - x.double() @ w.double()).abs().max())
print(f"contraction 32768 x 4096 x 64: {s.elapsed_time(e) / 10:.3f} ms per call (splits included), max abs err {err:.2e}")
