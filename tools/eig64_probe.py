#!/usr/bin/env python
"""Time / sweeps of the all-in-one n <= 64 solver on the 64 x 64 Gram matrices of a 512^3 phantom sweep
(second bond: first two sites against the rest; last-but-one bond: everything against the last two sites)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from bench import synthetic_volume                     # noqa: E402
from imgcompressionmps import _native, _ops            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vol = torch.from_numpy(synthetic_volume((n, n, n), 2027)).cuda()
dense = _ops.encode(vol).double()
ctx = _native.context()
cases = {"front (rows = first two sites)": dense.reshape(64, -1), "back (rows = last two sites)": dense.reshape(-1, 64).T.contiguous()}
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(64, 256, dtype=torch.float64, device="cuda", generator=g)
cases["random 64 x 256"] = a
for name, m in cases.items():
    gm = m @ m.T
    gm = 0.5 * (gm + gm.T)
    for opt in sys.argv[2:] or ["default"]:
        if "=" in opt:
            k, v = opt.split("=")
            ctx.set_option(k, int(v))
        ev, vec, sw = _ops.eigh(gm)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            _ops.eigh(gm)
        e.record()
        torch.cuda.synchronize()
        ref = torch.linalg.eigvalsh(gm).flip(0)
        err = float(((ev - ref).abs() / ref[0]).max())
        res = float(((gm @ vec - vec * ev).norm(dim=0) / ref[0]).max())
        orth = float((vec.T @ vec - torch.eye(64, dtype=torch.float64, device="cuda")).abs().max())
        print(f"{name} [{opt}]: {s.elapsed_time(e) / 5 * 1e3:.0f} us, {sw} sweeps, cond {float(ref[0] / ref[-1]):.2e}, "
              f"|dlam|/lam1 {err:.1e} resid {res:.1e} orth {orth:.1e}", flush=True)
