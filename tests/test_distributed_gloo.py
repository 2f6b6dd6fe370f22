"""Host-side sharding logic over a real 2-process gloo group on CPU (no GPU): item ownership,
gather order, failure on lost items.  The per-item work is a stand-in here - the device path
itself is covered by the -m gpu tests."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_dir):
    for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from imgcompressionmps.distributed import run_sharded, shard_indices
    from oracle.ndmps import OracleNDMPS
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    items = [rng.random((8, 8, 8)) for _ in range(n_items)]
    owned = []

    def work(i, vol):                       # stand-in for the device path: the oracle on tiny volumes
        owned.append(i)
        o = OracleNDMPS.from_tensor(vol, max_bond=4)
        return {"index": i, "bond_dims": o.bond_sizes(), "norm": float(o.norm_value)}

    res = run_sharded(items, work)
    assert owned == shard_indices(n_items, rank, world)
    assert [r["index"] for r in res] == list(range(n_items))
    np.save(Path(out_dir) / f"rank{rank}.npy", np.array([r["norm"] for r in res]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices():
    from imgcompressionmps.distributed import shard_indices
    assert shard_indices(7, 0, 2) == [0, 2, 4, 6] and shard_indices(7, 1, 2) == [1, 3, 5]
    assert shard_indices(3, 2, 8) == [2] and shard_indices(3, 5, 8) == []
    covered = sorted(i for r in range(8) for i in shard_indices(64, r, 8))
    assert covered == list(range(64))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_two_rank_gloo_gather(tmp_path):
    world, n_items = 2, 5
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert a.shape == (n_items,) and np.array_equal(a, b)       # every rank sees every result, same order
    # equals the single-process run
    sys.path.insert(0, str(ROOT))
    from oracle.ndmps import OracleNDMPS
    rng = np.random.default_rng(0)
    want = [float(OracleNDMPS.from_tensor(rng.random((8, 8, 8)), max_bond=4).norm_value) for _ in range(n_items)]
    assert np.allclose(a, want, rtol=1e-12)
