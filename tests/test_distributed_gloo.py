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


# ---- one tensor over several ranks: partition + the column-sharded sweep, on CPU --------------------
@pytest.mark.parametrize("shape,world", [((8, 12, 4), 2), ((8, 12, 4), 4), ((16, 16, 16), 8), ((6, 9), 3), ((6, 10, 15), 1)])
def test_shard_volume_is_a_block_of_the_last_site_index(shape, world):
    """shard_volume cuts exactly the voxels whose last site index lies in the rank's block, and
    encoding the sub-lattice with the local factor table gives dense[..., block]."""
    from imgcompressionmps.distributed import last_level_split, local_factors, place_shard, shard_volume
    from imgcompressionmps.utils.core import gen_encoding_map, get_factorlist
    from oracle import encoding as OE
    factors, _ = get_factorlist(shape)
    dims, enc = gen_encoding_map(shape)
    rng = np.random.default_rng(1)
    x = rng.random(shape)
    dense = OE.encode(x)
    dl = int(dims[-1]) // world
    full = np.zeros(shape)
    for r in range(world):
        part = shard_volume(x, factors, r, world)
        last = shard_volume(enc[-1], factors, r, world)
        assert last.min() >= r * dl and last.max() < (r + 1) * dl
        lf = local_factors(factors, world)
        assert tuple(part.shape) == tuple(n // w for n, w in zip(shape, last_level_split(factors, world)))
        # local encode: mixed-radix digits of the local coordinates with the local factors
        ldims = [int(d) for d in np.prod(lf, axis=1)]
        local_dense = _encode_with_factors(part, lf)
        assert np.array_equal(local_dense.reshape(ldims), dense.reshape([int(d) for d in dims])[..., r * dl:(r + 1) * dl])
        place_shard(full, part, factors, r, world)
    assert np.array_equal(full, x)


def test_last_level_split_rejects_bad_worlds():
    from imgcompressionmps.distributed import last_level_split
    from imgcompressionmps.utils.core import get_factorlist
    factors, _ = get_factorlist((8, 12, 4))               # last level (4, 3, 2)
    assert [int(f) for f in factors[-1]] == [4, 3, 2]
    assert last_level_split(factors, 2) == [2, 1, 1]
    assert last_level_split(factors, 4) == [4, 1, 1]
    assert last_level_split(factors, 12) == [4, 3, 1]
    assert last_level_split(factors, 24) == [4, 3, 2]
    for bad in (3, 5, 8, 48):
        with pytest.raises(ValueError):
            last_level_split(factors, bad)


def _encode_with_factors(x, factors):
    """Plain numpy statement of the encoding for an explicit factor table (utils/core.py:6-35)."""
    levels, ndim = factors.shape
    site = np.zeros((levels,) + x.shape, dtype=np.int64)
    for a in range(ndim):
        coord = np.arange(x.shape[a], dtype=np.int64)
        view = [1] * ndim
        view[a] = x.shape[a]
        below = 1
        digits = [None] * levels
        for lvl in range(levels - 1, -1, -1):
            digits[lvl] = (coord // below) % int(factors[lvl, a])
            below *= int(factors[lvl, a])
        for lvl in range(levels):
            site[lvl] = site[lvl] * int(factors[lvl, a]) + digits[lvl].reshape(view)
    dims = [int(d) for d in np.prod(factors, axis=1)]
    flat = np.zeros(x.shape, dtype=np.int64)
    for lvl in range(levels):
        flat = flat * dims[lvl] + site[lvl]
    out = np.empty(x.size, dtype=x.dtype)
    out[flat.reshape(-1)] = x.reshape(-1)
    return out


def _sharded_sweep_worker(rank, world, port, out_dir):
    """numpy emulation of ndmps_ttsvd_sharded with a real gloo allreduce: local Gram, allreduce,
    identical eigensolve, local projection; then gather + interleave + replicated tail."""
    for p in (str(ROOT), str(ROOT / "img-compression-mps_b200")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from imgcompressionmps.distributed import local_factors, shard_volume
    from imgcompressionmps.utils.core import get_factorlist
    from oracle import mps as OMPS
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape, chi = (16, 16, 16), 6
    rng = np.random.default_rng(5)
    x = rng.random(shape) + np.add.outer(np.add.outer(np.arange(16.0), np.arange(16.0)), np.arange(16.0)) / 16
    factors, _ = get_factorlist(shape)
    dims = [int(d) for d in np.prod(factors, axis=1)]
    lf = local_factors(factors, world)
    ldims = [int(d) for d in np.prod(lf, axis=1)]
    m = _encode_with_factors(shard_volume(x, factors, rank, world), lf)
    cores, r_prev, stop_site = [], 1, len(dims) - 2
    for i in range(stop_site):
        m = m.reshape(r_prev * ldims[i], -1)
        g = torch.from_numpy(m @ m.T)
        dist.all_reduce(g)                                  # the one collective of a sharded step
        lam, u = np.linalg.eigh(g.numpy())
        lam, u = lam[::-1], u[:, ::-1]
        s = np.sqrt(np.clip(lam, 0, None))
        n = OMPS.n_keep(s, 1e-10, "rsum2", chi)
        f = OMPS.renorm_factor(s, n, 2)
        cores.append((u[:, :n] * 1.0).reshape((dims[0], n) if i == 0 else (r_prev, dims[i], n)))
        m = f * (u[:, :n].T @ m)
        r_prev = n
    parts = [torch.zeros(m.shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(m)))
    dl = ldims[-1]
    gathered = np.stack([p.numpy() for p in parts]).reshape(world, r_prev, -1, dl)
    rem = gathered.transpose(1, 2, 0, 3).reshape(r_prev, -1)          # what ndmps_interleave_shards does
    tail = OMPS.tt_svd(rem.reshape(-1), [r_prev * dims[stop_site]] + dims[stop_site + 1:], max_bond=chi)
    tail[0] = tail[0].reshape(r_prev, dims[stop_site], -1)
    cores += tail
    want = OMPS.tt_svd(_encode_with_factors(x, factors), dims, max_bond=chi)
    assert OMPS.bond_sizes(cores) == OMPS.bond_sizes(want)
    rec, ref = OMPS.contract_dense(cores), OMPS.contract_dense(want)
    assert np.linalg.norm(rec - ref) / np.linalg.norm(ref) < 1e-9
    np.save(Path(out_dir) / f"bonds{rank}.npy", np.array(OMPS.bond_sizes(cores)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_sweep(tmp_path):
    mp.spawn(_sharded_sweep_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert np.array_equal(np.load(tmp_path / "bonds0.npy"), np.load(tmp_path / "bonds1.npy"))


def test_shard_partition_random_shapes():
    """Every voxel lands in exactly one rank's sub-lattice, for every world size the last level admits."""
    from imgcompressionmps.distributed import last_level_split, place_shard, shard_volume
    from imgcompressionmps.utils.core import get_factorlist
    rng = np.random.default_rng(11)
    primes = [2, 2, 2, 3, 5]
    for _ in range(12):
        ndim = int(rng.integers(1, 5))
        shape = tuple(int(np.prod(rng.choice(primes, size=int(rng.integers(1, 4))))) for _ in range(ndim))
        factors, _ = get_factorlist(shape)
        d_last = int(np.prod(factors[-1]))
        x = rng.random(shape)
        for world in range(1, d_last + 1):
            try:
                last_level_split(factors, world)
            except ValueError:
                continue
            seen = np.zeros(shape, dtype=np.int64)
            full = np.zeros(shape)
            for r in range(world):
                place_shard(seen, shard_volume(np.ones(shape, dtype=np.int64), factors, r, world) + shard_volume(seen, factors, r, world),
                            factors, r, world)
                place_shard(full, shard_volume(x, factors, r, world), factors, r, world)
            assert np.array_equal(seen, np.ones(shape, dtype=np.int64)), (shape, world)
            assert np.array_equal(full, x)
