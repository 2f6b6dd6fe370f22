"""Multi-GPU partitioning of the NDMPS path (one process per GPU, ``torch.distributed``).

The reference processes lists of tensors one after the other (``evaluation/benchmark.py:73-76,
117-118``); items are independent, so they shard across ranks with NO data-path collective:
item ``i`` goes to rank ``i % world`` and every rank runs the whole single-GPU path.  The only
communication is a gather of the per-item scalars (metrics, bond dimensions) at the end.

``process_group`` may be NCCL (GPU ranks) or gloo (the CPU tests of the host logic, where the
``compress_fn`` is a stand-in).

ONE very large tensor shards too: rank g keeps the voxels whose
last-level digits select block g of the LAST site index -- a sub-lattice of the volume, cut at load
time (``shard_volume``) -- which is a column block of every unfolding of the sweep.  The left
factor of a step depends only on ``G = M M^T``, a sum over column blocks, so each step is: local
Gram, allreduce of the bond-sized Gram matrix (``ShardedNDMPS``; NCCL through
``torch.distributed`` on the compute stream), the same small eigenproblem on every rank, local
projection.  Once the remainder is a few MB it is all-gathered and the sweep finishes replicated.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: item i belongs to rank i % world."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


def run_sharded(items: Sequence, compress_fn: Callable[[int, object], Dict], process_group=None) -> List[Dict]:
    """Run ``compress_fn(index, item)`` on this rank's items and gather the (small, picklable)
    result dicts of all ranks, returned in item order on every rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
    else:
        rank, world = 0, 1
    mine = [(i, compress_fn(i, items[i])) for i in shard_indices(len(items), rank, world)]
    if world == 1:
        return [r for _, r in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=process_group)
    merged = {i: r for part in gathered for i, r in part}
    if sorted(merged) != list(range(len(items))):
        raise RuntimeError("sharded run lost or duplicated items")
    return [merged[i] for i in range(len(items))]


def compress_and_score(index: int, volume, max_bond=None, cutoff: float = 1e-10, mode: str = "Std") -> Dict:
    """The per-item work of the reference's benchmark loop on the device: encode + truncate +
    reconstruct, then the metrics (``evaluation/benchmark.py:121-146``)."""
    from .core.ndmps import NDMPS
    from .utils.metrics import compute_overlap, compute_psnr, compute_ssim_by_dim
    full = NDMPS.from_tensor(volume, mode=mode, cutoff=cutoff)
    obj = NDMPS.from_tensor(volume, mode=mode, cutoff=cutoff, max_bond=max_bond)
    rec = obj.to_tensor_device()
    out = {"index": index, "bond_dims": obj.bond_sizes(), "compression_ratio": float(obj.compression_ratio()),
           "psnr": float(compute_psnr(rec, volume)), "fidelity": float(compute_overlap(obj, full))}
    if 2 <= len(obj.shape) <= 4 and min(obj.shape[:3]) >= 3:
        out["ssim"] = float(compute_ssim_by_dim(rec, volume))
    return out


# ---------------------------------------------------------------------------------------------
# one tensor over several ranks: column-sharded sweep
# ---------------------------------------------------------------------------------------------
def last_level_split(factors: np.ndarray, world: int) -> List[int]:
    """How ``world`` ranks divide the last site index: per-axis counts ``w`` with prod(w) == world.

    The last site index is the mixed-radix number of the axes' last-level digits, axis 0 most
    significant (``utils/core.py:6-35``).  Consecutive blocks of it are products of per-axis digit
    ranges only if the leading axes are split completely before the next one is touched."""
    last = [int(f) for f in np.asarray(factors)[-1]]
    if world < 1:
        raise ValueError("world must be >= 1")
    split, left = [], int(world)
    for f in last:
        if left == 1:
            split.append(1)
        elif left % f == 0:
            split.append(f)
            left //= f
        elif f % left == 0:
            split.append(left)
            left = 1
        else:
            raise ValueError(f"{world} ranks do not divide the last-level factors {last} block-wise")
    if left != 1:
        raise ValueError(f"{world} ranks exceed the last site dimension {int(np.prod(last))}")
    return split


@lru_cache(maxsize=32)
def _plan_cached(shape: Tuple[int, ...], factor_bytes: bytes, levels: int):
    from . import _native
    return _native.Plan(shape, np.frombuffer(factor_bytes, dtype=np.int64).reshape(levels, len(shape)).copy())


def local_plan(shape: Sequence[int], factors: np.ndarray, world: int):
    """(plan, local shape) of a rank's sub-lattice; plans are cached (building the tile tables and
    uploading them is milliseconds of host work and a synchronising copy)."""
    lf = np.ascontiguousarray(local_factors(factors, world), dtype=np.int64)
    lshape = tuple(int(n) // w for n, w in zip(shape, last_level_split(factors, world)))
    return _plan_cached(lshape, lf.tobytes(), int(lf.shape[0])), lshape


def _digit_ranges(factors: np.ndarray, rank: int, world: int) -> List[Tuple[int, int]]:
    split = last_level_split(factors, world)
    if not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    last = [int(f) for f in np.asarray(factors)[-1]]
    coords, g = [], int(rank)
    for w in reversed(split):                       # rank = mixed radix over the split, axis 0 most significant
        coords.append(g % w)
        g //= w
    coords.reverse()
    return [(c * (f // w), (c + 1) * (f // w)) for c, f, w in zip(coords, last, split)]


def local_factors(factors: np.ndarray, world: int) -> np.ndarray:
    """Factor table of a rank's sub-lattice: the last level divided by the split."""
    out = np.array(factors, dtype=np.int64, copy=True)
    out[-1] //= np.asarray(last_level_split(factors, world), dtype=np.int64)
    return out


def _lattice_view(array, factors: np.ndarray):
    """View with every axis a split into (n_a / f_a, f_a), f_a the last-level factor."""
    last = [int(f) for f in np.asarray(factors)[-1]]
    shape = []
    for n, f in zip(array.shape, last):
        shape += [int(n) // f, f]
    return array.reshape(shape)


def shard_volume(volume, factors: np.ndarray, rank: int, world: int):
    """This rank's sub-lattice of ``volume`` (numpy array or torch tensor, any device): the voxels
    whose last-level digits fall in the rank's block of the last site index.  Copy, C-contiguous."""
    ranges = _digit_ranges(factors, rank, world)
    view = _lattice_view(volume, factors)
    index = []
    for lo, hi in ranges:
        index += [slice(None), slice(lo, hi)]
    part = view[tuple(index)]
    shape = [part.shape[2 * a] * part.shape[2 * a + 1] for a in range(len(ranges))]
    return part.reshape(shape)


def place_shard(full, part, factors: np.ndarray, rank: int, world: int):
    """Inverse of ``shard_volume``: write a rank's sub-lattice back into the full array (in place)."""
    ranges = _digit_ranges(factors, rank, world)
    view = _lattice_view(full, factors)
    index, pshape = [], []
    for a, (lo, hi) in enumerate(ranges):
        index += [slice(None), slice(lo, hi)]
        pshape += [view.shape[2 * a], hi - lo]
    view[tuple(index)] = part.reshape(pshape)
    return full


class ShardedNDMPS:
    """MPS of ONE tensor whose voxels are spread over the ranks of a process group.

    ``from_local`` runs the sweep of ``NDMPS.from_tensor`` (``core/ndmps.py:36-78``, Std mode) on this
    rank's sub-lattice; the cores come out identical on every rank (and equal to what a single GPU
    computes on the whole tensor, up to the summation order of the Gram allreduce).
    ``to_local_tensor_device`` reconstructs this rank's sub-lattice (``core/ndmps.py:131-153``)."""

    # data-path collective accounting (read by bench.py): calls, payload bytes and device time of the Gram allreduces
    _ar = {"calls": 0, "bytes": 0, "events": []}

    @classmethod
    def allreduce_stats(cls, reset: bool = False) -> Dict:
        """{calls, bytes, ms} of the Gram allreduces since the last reset (ms from CUDA events on the compute
        stream; synchronises)."""
        ms = 0.0
        for beg, end in cls._ar["events"]:
            end.synchronize()
            ms += beg.elapsed_time(end)
        out = {"calls": cls._ar["calls"], "bytes": cls._ar["bytes"], "ms": ms}
        if reset:
            cls._ar = {"calls": 0, "bytes": 0, "events": []}
        return out

    def __init__(self, cores, site_dims, shape, factors, rank, world, singular_values):
        self.cores, self.site_dims, self.shape = cores, [int(d) for d in site_dims], tuple(shape)
        self.factors, self.rank, self.world = factors, int(rank), int(world)
        self.singular_values = singular_values

    def bond_sizes(self) -> List[int]:
        return [int(c.shape[-1]) for c in self.cores[:-1]]

    @classmethod
    def from_local(cls, local_volume, shape: Sequence[int], rank: Optional[int] = None, world: Optional[int] = None,
                   process_group=None, max_bond=None, cutoff: float = 1e-10, stop_bytes: int = 8 << 20) -> "ShardedNDMPS":
        import torch
        import torch.distributed as dist
        from . import _native, _ops
        from .utils.core import get_factorlist
        if rank is None or world is None:
            rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
        shape = tuple(int(s) for s in shape)
        factors, _ = get_factorlist(shape)
        site_dims = [int(d) for d in np.prod(factors, axis=1)]
        plan, lshape = local_plan(shape, factors, world)
        if tuple(local_volume.shape) != lshape:
            raise ValueError(f"rank {rank}: local volume has shape {tuple(local_volume.shape)}, expected {lshape}")
        dense = _ops.encode(local_volume.contiguous(), 1.0, plan=plan)
        ldims = [int(d) for d in plan.site_dims]
        L = len(ldims)

        def allreduce(t):
            ar = ShardedNDMPS._ar
            timed = t.is_cuda and len(ar["events"]) < 4096
            if timed:
                beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                beg.record()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
            if timed:
                end.record()
                ar["events"].append((beg, end))
            ar["calls"] += 1
            ar["bytes"] += t.numel() * t.element_size()

        cores, ranks, svals, rem = _ops.ttsvd_sharded(dense, ldims, world, allreduce if world > 1 else None,
                                                      stop_bytes=stop_bytes, cutoff=cutoff, max_bond=max_bond)
        del dense
        done = len(cores)
        rows, cols = int(rem.shape[0]), int(rem.shape[1])
        if world > 1:                                   # remainders of all ranks, back in site order
            gathered = torch.empty((world,) + tuple(rem.shape), dtype=rem.dtype, device=rem.device)
            if dist.get_backend(process_group) == "nccl":
                dist.all_gather_into_tensor(gathered, rem.contiguous(), group=process_group)
            else:                                       # gloo (tests): list form
                parts = [torch.empty_like(rem) for _ in range(world)]
                dist.all_gather(parts, rem.contiguous(), group=process_group)
                for g, part in enumerate(parts):
                    gathered[g].copy_(part)
            dl = ldims[-1]
            rem = _ops.interleave_shards(gathered, world, rows, cols // dl, dl)
        # replicated tail: the remainder as a dense array whose first site carries the incoming bond
        tail_dims = [rows * site_dims[done]] + site_dims[done + 1:]
        tcores, tranks, tsv = _ops.ttsvd(rem.reshape(-1), tail_dims, cutoff=cutoff, max_bond=max_bond)
        if len(tail_dims) == 1:
            tcores = [tcores[0].view(rows, site_dims[done])]
        else:
            first = tcores[0].view(rows, site_dims[done], tranks[0]) if done > 0 else tcores[0]
            tcores = [first] + list(tcores[1:])
        return cls(cores + tcores, site_dims, shape, factors, rank, world, list(svals) + list(tsv))

    def to_local_tensor_device(self):
        """This rank's sub-lattice of the reconstruction (device tensor)."""
        from . import _ops
        plan, lshape = local_plan(self.shape, self.factors, self.world)
        dl = int(plan.site_dims[-1])
        last = self.cores[-1]
        block = last[..., self.rank * dl:(self.rank + 1) * dl].contiguous()
        dense = _ops.contract_dense(list(self.cores[:-1]) + [block])
        return _ops.decode(dense.reshape(-1), lshape, plan=plan)
