#!/usr/bin/env python
"""K1 permutation at the BASELINE power-of-two shapes: register bit-permutation (permute_path 0) against the
shared-memory tiles (1): bit-exact agreement and CUDA-event times, algorithmic 8 B/voxel against the measured copy peak.
Between timed launches a 512 MB buffer is rewritten so that a 256^3 volume (67 MB) does not stay in the 126 MB L2."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from imgcompressionmps import _native, _ops            # noqa: E402

peak = 6559.0
try:
    peak = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
ctx = _native.context()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(256, 256, 256), (512, 512, 512), (2048, 2048)]
for shape in shapes:
    g = torch.Generator(device="cuda").manual_seed(1)
    vol = torch.randn(shape, dtype=torch.float32, device="cuda", generator=g)
    ref = None
    for path in (1, 0):
        ctx.set_option("permute_path", path)
        dense = _ops.encode(vol)
        back = _ops.decode(dense, shape)
        assert torch.equal(back, vol), ("decode", shape, path)
        if ref is None:
            ref = dense
        else:
            assert torch.equal(dense, ref), ("encode", shape, path)
        times = {"encode": [], "decode": []}
        for _ in range(7):
            for name in times:
                flush.fill_(1)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                if name == "encode":
                    _ops.encode(vol)
                else:
                    _ops.decode(dense, shape)
                e.record()
                torch.cuda.synchronize()
                times[name].append(s.elapsed_time(e))
        for name, t in times.items():
            t = sorted(t)[len(t) // 2]                 # includes the output allocation (cached) and the launch
            gbps = 8.0 * vol.numel() / t / 1e6
            print(f"{'x'.join(map(str, shape))} path {path} {name}: {t * 1e3:.1f} us  {gbps:.0f} GB/s  {gbps / peak:.3f} of {peak:.0f}",
                  flush=True)
    ctx.set_option("permute_path", 0)
print("permute probe ok")
