"""Env sharding + dataset gather: world_size-2 gloo on CPU (the N>1 host logic of bench.py /
SOARM101DataGenerator).  Shards are produced by the CPU oracle here; on GPUs the same code path
runs with NCCL and the CUDA rollout."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lerobot_mujoco_sim2real_b200 import sharding


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 4096, 1048576, 1048577):
        for w in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = sharding.shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, n_total, T, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lerobot_mujoco_sim2real_b200 import builtin_tables
        from oracle import oracle as O
        tables = builtin_tables()
        lo, hi = sharding.shard_range(n_total, rank, world)
        rows, _, _ = O.rollout(tables, O.make_spec(kind=0, seed=42, env_offset=lo), hi - lo, T, 10, nthreads=2)
        full = sharding.gather_rows(torch.from_numpy(rows), n_total, dst=0)
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_gather_is_invariant_to_world_size(tmp_path, oracle_mod, tables_v):
    n_total, T = 13, 3     # odd: ragged shards (7 + 6)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "rows.npy")
    mp.spawn(_worker, args=(2, port, n_total, T, out), nprocs=2, join=True)
    got = np.load(out)
    ref, _, _ = oracle_mod.rollout(tables_v, oracle_mod.make_spec(kind=0, seed=42), n_total, T, 10)
    assert got.shape == (n_total, T + 1, 13)
    np.testing.assert_array_equal(got, ref)    # global env ids key the RNG: bit-identical for any G


def _worker_edge(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        for n_total in (1, 2, 5):                       # 1 < world: rank 1's shard is EMPTY and must not hang the gather
            lo, hi = sharding.shard_range(n_total, rank, world)
            rows = (torch.arange(lo, hi, dtype=torch.float64)[:, None, None] * 100
                    + torch.arange(4, dtype=torch.float64)[None, :, None] * 10 + torch.arange(13, dtype=torch.float64))
            flags = torch.arange(lo, hi, dtype=torch.int32) * 3
            out = torch.full((n_total, 4, 13), -1.0, dtype=torch.float64) if rank == 0 else None
            got = sharding.gather_rows(rows, n_total, dst=0, out=out)       # received in place, no torch.cat
            gf = sharding.gather_rows(flags, n_total, dst=0)
            if rank == 0:
                assert got is out
                res[n_total] = (got.numpy().copy(), gf.numpy().copy())
            else:
                assert got is None and gf is None
        if rank == 0:
            np.savez(out_path, **{f"rows_{k}": v[0] for k, v in res.items()}, **{f"flags_{k}": v[1] for k, v in res.items()})
    finally:
        dist.destroy_process_group()


def test_gather_ragged_empty_and_in_place(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "edge.npz")
    mp.spawn(_worker_edge, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    for n_total in (1, 2, 5):
        want = (np.arange(n_total)[:, None, None] * 100.0 + np.arange(4)[None, :, None] * 10.0 + np.arange(13.0))
        np.testing.assert_array_equal(got[f"rows_{n_total}"], want)
        np.testing.assert_array_equal(got[f"flags_{n_total}"], np.arange(n_total) * 3)


def test_single_process_gather_is_identity():
    x = torch.arange(12.0).reshape(4, 3)
    assert sharding.gather_rows(x, 4) is x
    assert sharding.dist_info() == (0, 1)
