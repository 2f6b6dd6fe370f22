"""Concurrent processing of independent tensors on ONE GPU.

The reference's benchmark walks a list of tensors one after the other
(``evaluation/benchmark.py:73-76,117-118``).  On the device a single tensor cannot fill the
machine: the bond-sized eigenproblems of the sweep (``core/ndmps.py:74``) are latency-bound
cooperative kernels on a fraction of the SMs, and the host waits for each truncation decision.
Items are independent, so ``VolumePipeline`` keeps several of them in flight: ``workers`` host
threads, each with its own CUDA stream and its own native context (workspace arena, pinned
scratch -- ``_native.context()`` is per thread), pull items from a queue.  ctypes releases the
GIL for the duration of every native call, so the threads really overlap; on the GPU the
latency-bound kernels of one item run beside the bandwidth-bound kernels of another.

Nothing here changes results: every item goes through exactly the calls a single-threaded
caller would make.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Iterable, List, Optional, Sequence


class VolumePipeline:
    """``workers`` host threads x (CUDA stream + native context) on the current device."""

    def __init__(self, workers: int = 3, device: Optional[int] = None, blocking_sync: Optional[bool] = None):
        import torch
        if workers < 1:
            raise ValueError("workers must be >= 1")
        if not torch.cuda.is_available():
            from . import _native
            raise _native.NativeError("VolumePipeline needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.workers = int(workers)
        # host threads that wait for the GPU spin by default; with more waiting threads (all ranks of this
        # node) than cores they would starve each other, so they sleep on blocking events instead
        if blocking_sync is None:
            ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
            blocking_sync = self.workers * ranks_here > max(1, (os.cpu_count() or 1) - 1)
        self.blocking_sync = bool(blocking_sync)
        self._tls = threading.local()
        self._ctx_lock = threading.Lock()
        self._contexts = []                             # the workers' native contexts (for launch counts / options)
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="ndmps-worker",
                                        initializer=self._init_worker)

    # -- worker side ---------------------------------------------------------------------------
    def _init_worker(self):
        import torch
        from . import _native
        torch.cuda.set_device(self.device)
        self._tls.stream = torch.cuda.Stream(device=self.device)
        self._tls.done = torch.cuda.Event(blocking=self.blocking_sync)
        ctx = _native.context()
        if self.blocking_sync:                          # sleep, do not spin, while the GPU works
            ctx.set_option("blocking_sync", 1)
        with self._ctx_lock:
            self._contexts.append(ctx)

    def _run(self, fn: Callable, item, ready_event):
        import torch
        stream = self._tls.stream
        with torch.cuda.stream(stream):
            if ready_event is not None:
                stream.wait_event(ready_event)          # inputs produced on the submitting stream
            out = fn(item)
            self._tls.done.record(stream)
            self._tls.done.synchronize()                # the result is complete when the future resolves
        return out

    # -- caller side ---------------------------------------------------------------------------
    def map(self, fn: Callable, items: Iterable) -> List:
        """``[fn(item) for item in items]`` with up to ``workers`` items in flight, in item order.
        ``fn`` runs with the worker's stream current, so everything it launches (native calls
        included) lands there; device inputs must already be materialised on the calling
        stream, which the workers wait for."""
        import torch
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        futures = [self._pool.submit(self._run, fn, item, ready) for item in items]
        return [f.result() for f in futures]

    def roundtrip(self, volumes: Sequence, max_bond: Optional[int] = None, cutoff: float = 1e-10, mode: str = "Std",
                  keep: bool = True) -> List:
        """``NDMPS.from_tensor(v, ...).to_tensor_device()`` for every device tensor in ``volumes``."""
        from .core.ndmps import NDMPS

        def one(v):
            obj = NDMPS.from_tensor(v, mode=mode, cutoff=cutoff, max_bond=max_bond)
            rec = obj.to_tensor_device()
            return (obj, rec) if keep else None

        return self.map(one, volumes)

    def roundtrip_host(self, sources: Sequence, outs: Sequence, max_bond: Optional[int] = None,
                       cutoff: float = 1e-10) -> List:
        """Host-buffer round trips (``_ops.roundtrip_host``): host -> device copy, encode, sweep,
        reconstruct, decode, device -> host copy, several items in flight so the copies of one
        overlap the sweep of another.  ``sources`` / ``outs``: numpy arrays (pinned for overlap)."""
        from . import _ops

        def one(pair):
            src, dst = pair
            return _ops.roundtrip_host(src, max_bond=max_bond, cutoff=cutoff, out=dst)

        return self.map(one, list(zip(sources, outs)))

    def launch_count(self) -> int:
        """Kernels launched so far by all workers of this pipeline."""
        with self._ctx_lock:
            return sum(c.launch_count() for c in self._contexts)

    def close(self):
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
