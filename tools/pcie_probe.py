#!/usr/bin/env python
"""Host link of this box: pinned H2D alone, D2H alone, both at once on two streams (512 MB buffers, CUDA events /
host clock around a synchronize).  Says what the `e2e` line of bench.py can reach: a 512^3 float32 volume is 537 MB
each way."""
import time

import torch

n = 512 << 20
dev = torch.device("cuda:0")
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(3)
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.full((n,), 5, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def both_chunked(chunk=64 << 20):
    for o in range(0, n, chunk):
        with torch.cuda.stream(s1):
            d_a[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[o:o + chunk].copy_(d_b[o:o + chunk], non_blocking=True)


gb = n / 1e9
t = timed(h2d)
print(f"H2D alone      : {gb / t:6.1f} GB/s")
t = timed(d2h)
print(f"D2H alone      : {gb / t:6.1f} GB/s")
t = timed(both)
print(f"both, 2 streams: {2 * gb / t:6.1f} GB/s total ({gb / t:.1f} per direction)")
t = timed(both_chunked)
print(f"both, 64 MB chunks: {2 * gb / t:6.1f} GB/s total")
# a busy GPU beside the copies (the sweep kernels of other volumes): does HBM / SM load change the link rate?
x = torch.empty(256 << 20, dtype=torch.float32, device=dev)
y = torch.empty_like(x)


def both_busy():
    both()
    for _ in range(12):
        y.copy_(x)


t = timed(both_busy)
print(f"both + device copies on the default stream: {2 * gb / t:6.1f} GB/s total")
try:
    import os
    print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
    for node in sorted(p for p in os.listdir("/sys/devices/system/node") if p.startswith("node")):
        print(node, open(f"/sys/devices/system/node/{node}/cpulist").read().strip())
except Exception as exc:          # noqa: BLE001
    print("topology:", exc)
