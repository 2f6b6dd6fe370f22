#!/usr/bin/env python
"""Which eigensolver every bond of the BASELINE workloads takes (NDMPS_OPT_VERBOSE lines of the library on stderr)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import bench                                            # noqa: E402
from imgcompressionmps import _native                  # noqa: E402

ctx = _native.context()
for wl in sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5b"]:
    host = bench.make_input(wl, 0, 0)
    x = torch.from_numpy(host).cuda()
    unit = bench.device_unit(wl, bench.WORKLOADS[wl]["chi"])
    unit(x)
    torch.cuda.synchronize()
    print(f"==== {wl} {tuple(x.shape)}", file=sys.stderr, flush=True)
    ctx.set_option("verbose", 1)
    unit(x)
    torch.cuda.synchronize()
    ctx.set_option("verbose", 0)
