#!/usr/bin/env python
"""Which tcgen05 stage costs reconstruction parity?  exact path vs (exact Gram + tcgen05 projection / contraction)."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    sys.path.insert(0, p)
from bench import synthetic_fmri, synthetic_video, synthetic_volume
from imgcompressionmps import _native as N
from imgcompressionmps.core.ndmps import NDMPS

ctx = N.context()
cases = [("video DCT", lambda: synthetic_video((1920, 1080, 64), 4000, channel=1, chunk=2), "DCT"),
         ("fMRI", lambda: synthetic_fmri((64, 64, 32, 400), 3000), "Std"),
         ("512^3", lambda: synthetic_volume((512, 512, 512), 2027), "Std")]
for name, gen, mode in cases:
    x = torch.from_numpy(gen()).cuda()
    ctx.set_option("tc", 0); ctx.set_option("gram_path", 0)
    ref = NDMPS.from_tensor(x, mode=mode, max_bond=64)
    r0 = ref.to_tensor_device()
    for label, tc, gp in (("exact Gram + tc projection/contraction", 1, 1), ("int8 five-digit Gram + tc projection/contraction", 1, 3)):
        ctx.set_option("tc", tc); ctx.set_option("gram_path", gp)
        o = NDMPS.from_tensor(x, mode=mode, max_bond=64)
        r = o.to_tensor_device()
        sv = max(float(np.abs(a - b).max() / b[0]) for a, b in zip(o.singular_values, ref.singular_values))
        rel = float(torch.linalg.vector_norm((r - r0).double()) / torch.linalg.vector_norm(r0.double()))
        print(f"{name}: {label}: bonds equal {o.bond_sizes() == ref.bond_sizes()} max dsigma/s1 {sv:.2e} rec rel diff {rel:.2e}", flush=True)
    del x, r0, ref
ctx.set_option("tc", 1); ctx.set_option("gram_path", 0)
