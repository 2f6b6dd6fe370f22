#!/usr/bin/env python
"""Leading-64 solves for 1024 < n <= 1536: register-resident tridiagonalisation (topk_mid = 1) against the L2-streaming
kernel (topk_mid = 0): time, eigenvalue error, orthogonality, residual."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from imgcompressionmps import _native, _ops   # noqa: E402

ctx = _native.context()
g = torch.Generator(device="cuda").manual_seed(11)
for n in (1025, 1100, 1280, 1535, 1536):
    a = torch.randn(n, 2 * n, dtype=torch.float64, device="cuda", generator=g) * torch.logspace(0, -3, n, dtype=torch.float64, device="cuda")[:, None]
    gm = a @ a.T
    gm = 0.5 * (gm + gm.T)
    ref = torch.linalg.eigvalsh(gm).flip(0)
    for mid in (0, 1):
        ctx.set_option("topk_mid", mid)
        ev, vec, tr, health = _ops.eigh_topk(gm, 64)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            _ops.eigh_topk(gm, 64)
        e.record()
        torch.cuda.synchronize()
        rel = float(((ev - ref[:64]).abs() / ref[0]).max())
        orth = float((vec.T @ vec - torch.eye(64, dtype=torch.float64, device="cuda")).abs().max())
        res = float(((gm @ vec - vec * ev).norm(dim=0) / ref[0]).max())
        print(f"n={n} topk_mid={mid}: {s.elapsed_time(e) / 3:.3f} ms health={health} |dlam|/lam1={rel:.1e} orth={orth:.1e} resid={res:.1e}", flush=True)
ctx.set_option("topk_mid", 1)
