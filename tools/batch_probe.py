"""Throughput of the device-resident round trip with several volumes in flight (VolumePipeline).
Run on the GPU box:  python tools/batch_probe.py [n] [volumes] [workers ...]"""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "img-compression-mps_b200"))
sys.path.insert(0, str(ROOT))
from bench import synthetic_volume  # noqa: E402
from imgcompressionmps.batch import VolumePipeline  # noqa: E402


def make_volume(shape, seed):
    if len(shape) == 3:
        return synthetic_volume(shape, seed)
    import numpy as np
    rng = np.random.default_rng(seed)                   # low-rank-ish structure + noise for other ranks
    ramp = np.linspace(0.0, 1.0, shape[-1], dtype=np.float32)
    return (0.1 * rng.random(shape, dtype=np.float32) + ramp).astype(np.float32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    shape = tuple(int(v) for v in os.environ["PROBE_SHAPE"].split(",")) if os.environ.get("PROBE_SHAPE") else (n, n, n)
    nvol = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    workers = [int(w) for w in sys.argv[3:]] or [1, 2, 3, 4]
    torch.cuda.set_device(0)
    vols = [torch.from_numpy(make_volume(shape, 2026 + i)).cuda() for i in range(min(nvol, 4))]
    vols = [vols[i % len(vols)] for i in range(nvol)]
    host = [torch.from_numpy(make_volume(shape, 2026)).pin_memory().numpy() for _ in range(min(nvol, 4))]
    outs = [torch.empty(shape, dtype=torch.float32).pin_memory().numpy() for _ in range(min(nvol, 4))]
    for w in workers:
        blocking = {"0": False, "1": True}.get(os.environ.get("PROBE_BLOCKING", ""))
        with VolumePipeline(workers=w, blocking_sync=blocking) as pipe:
            print(f"blocking_sync={pipe.blocking_sync}", end=" ")
            pipe.roundtrip(vols[:w], max_bond=64, keep=False)
            pipe.roundtrip(vols[:w], max_bond=64, keep=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe.roundtrip(vols, max_bond=64, keep=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            nvx = 1
            for v in shape:
                nvx *= v
            msg = f"{shape} chi=64, {nvol} volumes, {w} in flight: {dt * 1e3 / nvol:.2f} ms/volume, {nvol * nvx / dt / 1e9:.3f} Gvoxel/s"
            srcs = [host[i % len(host)] for i in range(nvol)]
            dsts = [outs[i % len(outs)] for i in range(nvol)] if w == 1 else None
            if dsts is None:    # distinct output buffers while several are in flight
                dsts = [outs[i % len(outs)] for i in range(nvol)]
            for _ in range(3):
                pipe.roundtrip_host(srcs[:2 * w], dsts[:2 * w], max_bond=64)
            t0 = time.perf_counter()
            pipe.roundtrip_host(srcs, dsts, max_bond=64)
            dt = time.perf_counter() - t0
            print(msg + f" | host buffers: {dt * 1e3 / nvol:.2f} ms/volume, {nvol * nvx / dt / 1e9:.3f} Gvoxel/s", flush=True)


if __name__ == "__main__":
    main()
