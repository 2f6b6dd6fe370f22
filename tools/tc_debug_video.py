#!/usr/bin/env python
"""Where does the tcgen05 path lose accuracy on the DCT-mode video chunk?  tc on/off, step by step."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    sys.path.insert(0, p)
from bench import synthetic_video
from imgcompressionmps import _native as N, _ops
from imgcompressionmps.core.ndmps import NDMPS

ctx = N.context()
x = torch.from_numpy(synthetic_video((1920, 1080, 64), 4000, channel=1, chunk=2)).cuda()
for mode in ("Std", "DCT"):
    res = {}
    for tc in (0, 1):
        ctx.set_option("tc", tc)
        o = NDMPS.from_tensor(x, mode=mode, max_bond=64)
        res[tc] = (o, o.to_tensor_device())
    (o0, r0), (o1, r1) = res[0], res[1]
    print(mode, "bonds", o0.bond_sizes(), o1.bond_sizes())
    for k, (a, b) in enumerate(zip(o1.singular_values, o0.singular_values)):
        print(f"  bond {k}: max |ds|/s1 {np.abs(a - b).max() / b[0]:.2e}  s_last/s1 {b[-1] / b[0]:.2e}")
    print("  rec rel diff", float(torch.linalg.vector_norm((r1 - r0).double()) / torch.linalg.vector_norm(r0.double())))
    # same cores, contraction only
    ctx.set_option("tc", 0); d0 = _ops.contract_dense(o0.mps.cores)
    ctx.set_option("tc", 1); d1 = _ops.contract_dense(o0.mps.cores)
    print("  contraction only rel diff", float(torch.linalg.vector_norm((d1 - d0).double()) / torch.linalg.vector_norm(d0.double())))
# the unfoldings of the DCT case
ctx.set_option("tc", 1)
v = _ops.dct_last_axis(x)
dense = _ops.encode(v)
print("dense dims", list(dense.shape))
m1 = dense.reshape(480, -1)
ctx.set_option("gram_path", 0); ctx.set_option("tc", 0); g0 = _ops.gram(m1)
ctx.set_option("gram_path", 3); g1 = _ops.gram(m1)
ctx.set_option("gram_path", 0); ctx.set_option("tc", 1)
print("gram 480 x 276480: max|dG|/max|G|", float((g1 - g0).abs().max() / g0.abs().max()), "diag rel", float(((g1 - g0).diagonal().abs() / g0.diagonal()).max()))
ev0 = torch.linalg.eigvalsh(g0).flip(0); ev1 = torch.linalg.eigvalsh(g1).flip(0)
print("  top-64 sigma diff / s1", float(((ev1[:64].clamp_min(0).sqrt() - ev0[:64].clamp_min(0).sqrt()).abs() / ev0[0].sqrt()).max()), "ev64/ev1", float(ev0[63] / ev0[0]))
# projection with a random isometry
q, _ = torch.linalg.qr(torch.randn(480, 64, dtype=torch.float64, device="cuda"))
want = q.T @ m1.double()
ctx.set_option("gemm_path", 3); ctx.set_option("gemm_out_t", 1)
c = torch.empty((64, m1.shape[1]), dtype=torch.float32, device="cuda")
mt = m1.T
N.check(N.load_library().ndmps_gemm(N.handle(), mt.shape[0], 64, 480, 1.0, N.ptr(m1), N.dtype_code(m1.dtype), mt.stride(0), mt.stride(1),
                                    N.ptr(q), N.dtype_code(q.dtype), q.stride(0), q.stride(1), N.ptr(c), N.dtype_code(c.dtype), mt.shape[0]), "gemm")
ctx.set_option("gemm_path", 0); ctx.set_option("gemm_out_t", 0)
print("projection 64 x 276480 (k = 480): max err / max", float((c.double() - want).abs().max() / want.abs().max()),
      " fro rel", float(torch.linalg.vector_norm(c.double() - want) / torch.linalg.vector_norm(want)))
# PCIe
h = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory(); d = torch.empty_like(h, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); [fn() for _ in range(4)]; e.record(); torch.cuda.synchronize()
    print(name, "512 MB pinned:", 4 * 0.536870912 / (s.elapsed_time(e) * 1e-3), "GB/s")
