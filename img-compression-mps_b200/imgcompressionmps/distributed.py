"""Multi-GPU partitioning of the NDMPS path (one process per GPU, ``torch.distributed``).

The reference processes lists of tensors one after the other (``evaluation/benchmark.py:73-76,
117-118``); items are independent, so they shard across ranks with NO data-path collective:
item ``i`` goes to rank ``i % world`` and every rank runs the whole single-GPU path.  The only
communication is a gather of the per-item scalars (metrics, bond dimensions) at the end.

``process_group`` may be NCCL (GPU ranks) or gloo (the CPU tests of the host logic, where the
``compress_fn`` is a stand-in).  Nothing in here touches the device itself.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence


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
