#!/usr/bin/env python
"""Headline benchmark: NDMPS encode + truncate-to-chi + reconstruct voxels/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2|cfg3] [--chi 64]

A step is one pass of the hot path over one BATCH of synthetic volumes (independent items, as
the reference's benchmark loop processes them): per volume ``NDMPS.from_tensor(vol,
max_bond=chi)`` (permute to site order, TT-SVD sweep, boundary list, norm) followed by
``to_tensor`` (contract, inverse permute); ``--in-flight`` volumes run concurrently on one
GPU (``imgcompressionmps.batch.VolumePipeline``: host threads x CUDA streams).  Default workload is
BASELINE.json configs[1]: a 256^3 float32 synthetic MRI volume at chi = 64; ``cfg3`` is
the 512^3 volume of the north-star target.

* ``value``  : device-resident - the volumes already sit in HBM, the reconstructions stay
  in HBM.  K steps timed with CUDA events, L2 flushed between steps (outside the event
  pairs), barrier + synchronize on both sides, max over ranks.
* ``single_volume`` : the same path one volume at a time (latency), with the stage profiler on;
  the per-kernel rooflines come from this pass.
* ``e2e``    : the same step through the C ABI on HOST buffers (``ndmps_roundtrip_host``:
  H2D copy, encode, sweep, contract, decode, D2H copy inside the timed region).
* N > 1     : one process per GPU (torchrun); every rank runs the whole path on its own
  volume - independent volumes shard with no data-path collective (weak scaling).
* ``--impl reference``: the CPU oracle (numpy/LAPACK float64 restatement of the reference
  path - the reference itself cannot be imported without quimb / scikit-image) on a bounded
  sample, all host threads, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "ndmps_encode_truncate_reconstruct_voxels_per_s"
UNIT = "voxels/s"
WORKLOADS = {
    "cfg2": {"shape": (256, 256, 256), "name": "configs[1]: 256x256x256 fp32 synthetic MRI volume", "in_flight": 12},
    "cfg3": {"shape": (512, 512, 512), "name": "configs[2]: 512x512x512 fp32 synthetic volume (north-star target)", "in_flight": 4},
}
CPU_SAMPLE_SHAPE = (128, 128, 128)


def synthetic_volume(shape, seed):
    """Ellipsoid phantom + smooth texture + 0.02 noise inside the object, exact-zero background
    (SURVEY section 8d).  Built per z-slab to keep host memory flat."""
    rng = np.random.default_rng(seed)
    params = [(rng.uniform(-0.25, 0.25, 3), rng.uniform(0.35, 0.8, 3) * (1.0 - 0.18 * k), 0.25 + 0.1 * k) for k in range(4)]
    axes = [np.linspace(-1, 1, n, dtype=np.float32) for n in shape]
    out = np.empty(shape, dtype=np.float32)
    gy, gz = np.meshgrid(axes[1], axes[2], indexing="ij")
    for i, xv in enumerate(axes[0]):
        sl = np.zeros(shape[1:], dtype=np.float32)
        for c, r, amp in params:
            r2 = ((xv - c[0]) / r[0]) ** 2 + ((gy - c[1]) / r[1]) ** 2 + ((gz - c[2]) / r[2]) ** 2
            sl += amp / (1.0 + np.exp(np.clip((r2 - 1.0) * 12.0, -60, 60)))
        inside = sl > 0.05
        sl += 0.05 * np.cos(3 * xv) * np.cos(3 * gy + 1) * np.cos(3 * gz + 2) * inside
        sl += 0.02 * rng.random(shape[1:], dtype=np.float32) * inside
        sl[sl < 0.02] = 0.0
        out[i] = sl
    return out


def synthetic_image(shape=(256, 256), seed=2025):
    """configs[0]: smooth + texture grayscale image, clip(sum of 3 Gaussians + 0.05 noise, 0, 1)."""
    rng = np.random.default_rng(seed)
    gy, gx = np.meshgrid(np.linspace(-1, 1, shape[0]), np.linspace(-1, 1, shape[1]), indexing="ij")
    img = np.zeros(shape)
    for _ in range(3):
        cy, cx = rng.uniform(-0.5, 0.5, 2)
        sy, sx = rng.uniform(0.15, 0.5, 2)
        img += rng.uniform(0.3, 0.6) * np.exp(-(((gy - cy) / sy) ** 2 + ((gx - cx) / sx) ** 2))
    img += 0.05 * rng.random(shape)
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def synthetic_fmri(shape=(64, 64, 32, 400), seed=3000):
    """configs[3]: static phantom x (1 + 0.05 low-rank temporal modes (rank 8, sinusoids)) + 0.01 noise."""
    rng = np.random.default_rng(seed)
    sx, sy, sz, nt = shape
    g = np.meshgrid(*[np.linspace(-1, 1, n, dtype=np.float32) for n in (sx, sy, sz)], indexing="ij")
    base = np.zeros((sx, sy, sz), dtype=np.float32)
    for k in range(3):
        c = rng.uniform(-0.2, 0.2, 3)
        r = rng.uniform(0.45, 0.85, 3) * (1.0 - 0.2 * k)
        r2 = sum(((gi - ci) / ri) ** 2 for gi, ci, ri in zip(g, c, r))
        base += (0.3 + 0.15 * k) / (1.0 + np.exp(np.clip((r2 - 1.0) * 10.0, -60, 60)))
    t = np.arange(nt, dtype=np.float32) / nt
    out = np.empty(shape, dtype=np.float32)
    mod = np.ones(shape, dtype=np.float32)
    for m in range(8):
        spatial = np.cos((m % 3 + 1) * g[0] + m) * np.cos((m % 2 + 1) * g[1] - m) * np.cos((m % 4) * 0.5 * g[2])
        temporal = np.sin(2 * np.pi * (m + 1) * t + rng.uniform(0, 6.28))
        mod += (0.05 / (1 + 0.5 * m)) * spatial[..., None].astype(np.float32) * temporal[None, None, None, :].astype(np.float32)
    np.multiply(base[..., None], mod, out=out)
    out += 0.01 * rng.random(shape, dtype=np.float32)
    return out


def synthetic_video(shape=(1920, 1080, 64), seed=4000, channel=0, chunk=0):
    """configs[4], non-degenerate partition 5-B: one colour channel x one 64-frame chunk of a moving-blob
    video: 5 Gaussians with linear motion, per-channel gain, 0.02 noise."""
    rng = np.random.default_rng(seed)
    blobs = [(rng.uniform(-0.6, 0.6, 2), rng.uniform(-0.4, 0.4, 2), rng.uniform(0.1, 0.35, 2), rng.uniform(0.3, 0.8)) for _ in range(5)]
    gains = rng.uniform(0.6, 1.0, 3)
    noise = np.random.default_rng(seed + 17 * channel + 101 * chunk + 1)
    w, h, nf = shape
    gx = np.linspace(-1, 1, w, dtype=np.float32)[:, None]
    gy = np.linspace(-1, 1, h, dtype=np.float32)[None, :]
    out = np.empty(shape, dtype=np.float32)
    for f in range(nf):
        tt = (chunk * nf + f) / 512.0
        fr = np.zeros((w, h), dtype=np.float32)
        for p0, vel, sig, amp in blobs:
            cx, cy = p0[0] + vel[0] * tt, p0[1] + vel[1] * tt
            fr += np.float32(amp) * (np.exp(-((gx - np.float32(cx)) / np.float32(sig[0])) ** 2) * np.exp(-((gy - np.float32(cy)) / np.float32(sig[1])) ** 2))
        fr *= np.float32(gains[channel])
        fr += 0.02 * noise.random((w, h), dtype=np.float32)
        out[:, :, f] = fr
    return out


def algorithmic_work(dims, ranks):
    """Bytes / flops of the stages per SURVEY section 8(d) for float32 payloads (4 B)."""
    L, n = len(dims), int(np.prod(dims))
    r = [1] + list(ranks) + [1]
    sweep_bytes = gram_flops = proj_flops = 0.0
    cols = n
    for i in range(L - 1):
        m = r[i] * dims[i]
        cols //= dims[i]
        sweep_bytes += 4.0 * (2 * m * cols + r[i + 1] * cols)
        gram_flops += 2.0 * m * m * cols
        proj_flops += 2.0 * m * r[i + 1] * cols
    recon_bytes = recon_flops = 0.0
    p = 1
    for k in range(L):
        recon_bytes += 4.0 * (p * r[k] + p * dims[k] * r[k + 1])
        recon_flops += 2.0 * p * r[k] * dims[k] * r[k + 1]
        p *= dims[k]
    # Gram passes actually executed: sites are front-merged while the fused row count stays <= 512
    executed = issued = 0.0
    cols, i, rprev = n, 0, 1
    while i < L - 1:
        rows = rprev * dims[i]
        cols //= dims[i]
        k = 1
        while i + k < L - 1 and rows * dims[i + k] <= 512 and rows * dims[i + k] <= cols // dims[i + k]:
            rows *= dims[i + k]
            cols //= dims[i + k]
            k += 1
        side = min(rows, cols)
        executed += 2.0 * side * side * max(rows, cols)
        nt = -(-side // 128)                                  # gram_dmma computes the upper-triangle 128 x 128 tiles only
        issued += 2.0 * side * side * max(rows, cols) * ((nt + 1) / (2.0 * nt) if side >= 48 else 1.0)
        rprev = r[i + k]
        i += k
    return {"encode_bytes": 8.0 * n, "decode_bytes": 8.0 * n, "sweep_bytes": sweep_bytes, "gram_flops": gram_flops,
            "gram_flops_executed": executed, "gram_flops_issued": issued,
            "project_flops": proj_flops, "recon_bytes": recon_bytes, "recon_flops": recon_flops}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def _all_host_threads():
    """Context manager giving BLAS every host core (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def cpu_oracle_rate(sample, chi, repeats):
    """voxels/s of the CPU oracle (numpy float64, LAPACK gesdd, all BLAS threads) on `sample`."""
    from oracle.ndmps import OracleNDMPS
    best = None
    with _all_host_threads():
        for _ in range(repeats):
            t0 = time.perf_counter()
            OracleNDMPS.from_tensor(sample, max_bond=chi).to_tensor()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return sample.size / best, best


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        with _all_host_threads():
            n = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    sample = synthetic_volume(CPU_SAMPLE_SHAPE, 2026)
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(sample, args.chi, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_rate(sample, args.chi, 1)
    dt = time.perf_counter() - t0
    value = sample.size * args.steps / dt
    cores = host_threads()
    sample_txt = (f"{'x'.join(map(str, CPU_SAMPLE_SHAPE))} float32 phantom of the same family per step (bounded sample "
                  f"of {wl['name']}), from_tensor(max_bond={args.chi}) + to_tensor, numpy/LAPACK float64")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "chi": args.chi, "mode": "Std", "sample": sample_txt},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference cannot be imported (quimb, scikit-image absent): oracle port timed on host cores",
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from imgcompressionmps import _native, _ops
    from imgcompressionmps.core.ndmps import NDMPS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for --impl ours")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from imgcompressionmps.batch import VolumePipeline

    wl = WORKLOADS[args.workload]
    shape = wl["shape"]
    nvox = int(np.prod(shape))
    in_flight = args.in_flight or wl["in_flight"]
    batch = args.volumes or 4 * in_flight                            # volumes per step (pipeline fill / drain amortised)
    n_distinct = min(batch, 4)                                       # distinct inputs (each >= 64 MB; L2 is flushed between steps)
    hosts = [synthetic_volume(shape, 2026 + 16 * rank + i) for i in range(n_distinct)]
    pinned = [torch.from_numpy(h).pin_memory() for h in hosts]
    vols_distinct = [p.cuda(non_blocking=False) for p in pinned]
    vols = [vols_distinct[i % n_distinct] for i in range(batch)]
    vol = vols[0]
    ctx = _native.context()
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    pipe = VolumePipeline(workers=in_flight)

    def single():
        obj = NDMPS.from_tensor(vol, max_bond=args.chi)
        rec = obj.to_tensor_device()
        return obj, rec

    def step():
        pipe.roundtrip(vols, max_bond=args.chi, keep=False)

    for _ in range(max(args.warmup, 3)):
        obj, rec = single()
        step()
    torch.cuda.synchronize()
    ranks = obj.bond_sizes()
    dims = obj.mps.site_dims
    err = float(torch.linalg.vector_norm((rec - vol).double()) / torch.linalg.vector_norm(vol.double()))

    # ---- device-resident timing: `batch` volumes per step, `in_flight` of them concurrently -------------
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = pipe.launch_count()
    barrier()
    with ClockSampler(local) as clocks:
        for i in range(args.steps):
            flush_buf.fill_(i & 0xFF)                                # evict L2 (outside the event pair)
            starts[i].record()
            step()                                                   # workers wait for the submitting stream, then are joined
            stops[i].record()
        barrier()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    launches = pipe.launch_count() - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * batch * nvox * args.steps / (total_ms * 1e-3)

    # ---- one volume at a time, stage profiler on: latency and the per-kernel rooflines ------------------
    p_starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    p_stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ctx.profile(True)
    ctx.stage_times(reset=True)
    ctx.stat("eig_flops", reset=True)
    barrier()
    for i in range(args.steps):
        flush_buf.fill_(i & 0xFF)
        p_starts[i].record()
        single()
        p_stops[i].record()
    barrier()
    single_ms = sum(s.elapsed_time(e) for s, e in zip(p_starts, p_stops)) / args.steps
    stages = ctx.stage_times(reset=True)
    ctx_eig_flops = ctx.stat("eig_flops", reset=True)
    ctx.profile(False)

    # ---- end to end through the C ABI on host buffers, same batch, same concurrency ----------------------
    srcs = [pinned[i % n_distinct].numpy() for i in range(batch)]
    n_dst = min(batch, 2 * in_flight)                                # outputs in flight never exceed in_flight: 2x is ample
    dst_t = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(n_dst)]
    dsts = [dst_t[i % n_dst].numpy() for i in range(batch)]
    for _ in range(3):
        pipe.roundtrip_host(srcs, dsts, max_bond=args.chi)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe.roundtrip_host(srcs, dsts, max_bond=args.chi)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * batch * nvox * args.steps / e2e_s
    src, dst = srcs[0], dsts[0]
    e2e_err = float(np.linalg.norm(dst.astype(np.float64) - src) / np.linalg.norm(src.astype(np.float64)))
    pipe.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the stages (algorithmic work / measured device time) ---------------------------
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        m = json.loads(pk.read_text())
        peaks = {"hbm_gbs": m["hbm_gbs"], "bf16_tflops": m.get("bf16_tflops_sustained", m["bf16_tflops"]), "source": "measured"}
    FP64_PEAK_TFLOPS = 34.2          # measured on this pool, tools/microbench/fp64_rate.cu (profiles/r01_summary.md)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture profiles/r01f_kernels_raw.csv (256^3 only)
    ncu_traffic = {"permute": 67.15e6 + 18.72e6, "gram": 67.13e6 + 5.30e6} if args.workload == "cfg2" else {}
    work = algorithmic_work(dims, ranks)
    per_step = {k: (v[0] / args.steps, v[1] / max(args.steps, 1)) for k, v in stages.items()}
    step_ms = total_ms / args.steps
    shares = {k: round(v[0] / single_ms, 4) for k, v in per_step.items() if v[1]}
    rooflines = {}
    gram_ms, gram_calls = per_step["gram"]
    if gram_calls:
        ach = work["gram_flops_executed"] / (gram_ms * 1e-3) / 1e12
        rooflines["gram"] = {
            "kernel": "gram_dmma_kernel (FP64 tensor pipe, mma.sync m8n8k4, exact float64 accumulation)", "bound": "tensor",
            "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
            "frac_of_fp64_peak": ach * work["gram_flops_issued"] / work["gram_flops_executed"] / FP64_PEAK_TFLOPS,
            "issued_tflops_fp64": ach * work["gram_flops_issued"] / work["gram_flops_executed"],
            "fp64_peak": FP64_PEAK_TFLOPS, "traffic": ncu_traffic.get("gram"),
            "ms_per_step": gram_ms, "calls_per_step": gram_calls,
            "flops_per_step": work["gram_flops_executed"],
            "note": "achieved = 2 m^2 C of the Gram passes actually run (front-merged group + later steps) / time; the kernel "
                    "issues only the upper-triangle tiles (issued_tflops_fp64).  tcgen05 has no float64 kind, so the bf16 peak is "
                    "the wrong denominator for this exact contraction - see frac_of_fp64_peak (issued / measured FP64 rate)"}
    perm_ms, perm_calls = per_step["permute"]
    if perm_calls:
        ach = (work["encode_bytes"] + work["decode_bytes"]) / (perm_ms * 1e-3) / 1e9
        rooflines["permute"] = {
            "kernel": "permute_tiled_kernel<float,4> (encode + decode)", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": ncu_traffic.get("permute"), "ms_per_step": perm_ms,
            "calls_per_step": perm_calls, "bytes_per_step": work["encode_bytes"] + work["decode_bytes"],
            "note": "achieved = algorithmic 8 B/voxel x 2 launches / event-timed stage (event overhead included; the kernel alone "
                    "is 35 us = 3.8 TB/s under ncu); traffic = DRAM bytes of ONE launch: below the algorithmic 134 MB because most "
                    "of the 64 MB output is still in the 126 MB L2 when the kernel ends"}
    eig_ms, eig_calls = per_step["eig"]
    eig_flops = ctx_eig_flops / args.steps
    if eig_calls:
        ach = eig_flops / (eig_ms * 1e-3) / 1e12
        rooflines["eig"] = {
            "kernel": "eig_topk.cu (tridiag_kernel, bisect, invit, Rayleigh-Ritz) for capped bonds; eig.cu Jacobi otherwise (float64 SIMT)",
            "bound": "latency (neither hbm nor tensor)",
            "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s (fp64)", "frac": ach / FP64_PEAK_TFLOPS, "traffic": None,
            "ms_per_step": eig_ms, "calls_per_step": eig_calls, "flops_per_step": eig_flops}
    dominant = max(shares, key=shares.get) if shares else None
    # headline roofline: the HBM-bound permutation (the path's pure data-movement kernel); the others ride along
    roof = dict(rooflines.get("permute") or rooflines.get("gram") or {})
    roof.update({"peak_source": peaks["source"], "dominant_stage_by_time": dominant, "stage_share_of_step": shares,
                 "all": rooflines})

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "shape": list(shape), "chi": args.chi, "mode": "Std", "site_dims": dims,
                   "bond_dims": ranks, "volumes_per_step_per_gpu": batch, "volumes_in_flight_per_gpu": in_flight,
                   "l2_flush_between_steps": True,
                   "parallelism": f"{world} GPU(s) x {batch} independent volumes per step, {in_flight} in flight per GPU "
                                  f"(host threads x CUDA streams), no data-path collective",
                   "reconstruction_rel_error_vs_input": err},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch * nvox * 4, "d2h_bytes_per_step": batch * nvox * 4,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "entry": "VolumePipeline.roundtrip_host -> ndmps_roundtrip_host (C ABI, pinned host buffers)",
                "reconstruction_rel_error_vs_input": e2e_err},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roof,
        "single_volume": {"ms": single_ms, "value": nvox / (single_ms * 1e-3), "unit": UNIT,
                          "note": "one volume at a time on one stream (latency); the stage times and rooflines are from this pass",
                          "stage_ms": {k: round(v[0], 4) for k, v in per_step.items() if v[1]}},
        "speed_of_light": {"hbm_floor_ms": (work["encode_bytes"] + work["decode_bytes"] + work["sweep_bytes"] + work["recon_bytes"])
                           / (peaks["hbm_gbs"] * 1e6), "note": "SURVEY 8(d) algorithmic bytes / measured HBM peak"},
    }
    if world == 1 and not args.no_cpu:
        sample = synthetic_volume(CPU_SAMPLE_SHAPE, 2026)
        rate, best = cpu_oracle_rate(sample, args.chi, 3)
        result["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": host_threads(), "kind": "port",
            "sample": f"{'x'.join(map(str, CPU_SAMPLE_SHAPE))} float32 phantom (bounded sample of the workload), "
                      f"from_tensor(max_bond={args.chi}) + to_tensor, numpy/LAPACK float64, best of 3 ({best:.2f} s)"}
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def run_sharded_volume(args):
    """ONE volume spread over the ranks (BASELINE configs[2]): every rank holds the
    sub-lattice of the volume that forms a column block of the unfoldings, the sweep allreduces the
    bond-sized Gram matrix of every step over NCCL, the reconstruction stays sharded.  Strong scaling."""
    import torch
    import torch.distributed as dist

    from imgcompressionmps import _native
    from imgcompressionmps.distributed import ShardedNDMPS, shard_volume
    from imgcompressionmps.utils.core import get_factorlist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = WORKLOADS[args.workload]
    shape = wl["shape"]
    nvox = int(np.prod(shape))
    factors, _ = get_factorlist(shape)
    host_local = np.ascontiguousarray(shard_volume(synthetic_volume(shape, 2027), factors, rank, world))   # cut at load time
    pinned = torch.from_numpy(host_local).pin_memory()
    out_pinned = torch.empty_like(pinned).pin_memory()
    vol = pinned.cuda()
    ctx = _native.context()
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step(v):
        obj = ShardedNDMPS.from_local(v, shape, rank=rank, world=world, max_bond=args.chi)
        return obj, obj.to_local_tensor_device()

    for _ in range(max(args.warmup, 3)):
        obj, rec = step(vol)
    err2 = torch.stack([((rec - vol).double() ** 2).sum(), (vol.double() ** 2).sum()])
    if world > 1:
        dist.all_reduce(err2)
    err = float(torch.sqrt(err2[0] / err2[1]))
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = ctx.launch_count()
    barrier()
    with ClockSampler(local) as clocks:
        for i in range(args.steps):
            flush_buf.fill_(i & 0xFF)
            starts[i].record()
            step(vol)
            stops[i].record()
        barrier()
    t = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, stops))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    launches = ctx.launch_count() - launches0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v = pinned.cuda(non_blocking=True)
        _, r = step(v)
        out_pinned.copy_(r, non_blocking=True)
        torch.cuda.synchronize()
    e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e.item())
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": nvox * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "shape": list(shape), "chi": args.chi, "mode": "Std", "bond_dims": obj.bond_sizes(),
                       "parallelism": f"ONE volume column-sharded over {world} GPU(s): local Gram + NCCL allreduce of the bond-sized "
                                      f"Gram matrix per sweep step, replicated eigensolve, local projection; sharded reconstruction",
                       "l2_flush_between_steps": True, "reconstruction_rel_error_vs_input": err},
            "e2e": {"value": nvox * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nvox * 4, "d2h_bytes_per_step": nvox * 4,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "entry": "ShardedNDMPS.from_local + to_local_tensor_device, pinned host shards"},
            "gpu_launches": int(launches) * world, "clocks": clocks.summary(),
            "roofline": None}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--chi", type=int, default=64)
    ap.add_argument("--in-flight", type=int, default=0, help="volumes processed concurrently per GPU (0: workload default)")
    ap.add_argument("--volumes", type=int, default=0, help="volumes per step per GPU (0: four times the number in flight)")
    ap.add_argument("--sharded", action="store_true",
                    help="ONE volume column-sharded over the GPUs (strong scaling) instead of independent volumes per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.sharded:
        run_sharded_volume(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
