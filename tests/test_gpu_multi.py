"""Two ranks, two GPUs, NCCL: the multi-GPU host path of SOARM101DataGenerator and sharding.SharedRows
(skipped on boxes with fewer than two GPUs; the gloo tests in test_sharding.py cover the host logic on CPU)."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _need_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs, this box has {torch.cuda.device_count()}")


def _args(d):
    return types.SimpleNamespace(
        xml_path="unused", x_dim=8, u_dim=5, device="cuda", seed=42, env="SOARM101",
        train_samples=301, train_steps=4, test_samples=51, test_steps=6, batch_size=32, eval_batch_size=16,
        data_dir_save=str(d), data_dir_load_train=os.path.join(str(d), "train_data_301_4.npy"),
        data_dir_load_val=os.path.join(str(d), "val_data_51_6.npy"))


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from lerobot_mujoco_sim2real_b200 import builtin_tables, sharding, tables as T
        from lerobot_mujoco_sim2real_b200.SOARM101_DataCollection import SOARM101DataGenerator
        from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
        tables = builtin_tables()
        # (1) SharedRows: both ranks' kernels write into rank 0's buffer; ragged shards (2049 + 2048)
        n_total, Tn = 4097, 7
        lo, hi = sharding.shard_range(n_total, rank, world)
        sr = sharding.SharedRows(n_total, (Tn + 1, T.ROW), torch.float64, rank, dst=0)
        env = SOARM101VecEnv(tables=tables, num_envs=hi - lo, device=rank)
        env.rollout(Tn, "chirp", seed=3, env_offset=lo, out_ptr=sr.local_ptr)
        full = sr.finish()
        if rank == 0:
            np.save(os.path.join(tmp, "shared.npy"), full.cpu().numpy())
        else:
            assert full is None
        del full
        sr.close()
        # (2) the data generator: sharded generation, flagged trajectories regenerated on rank 0, files visible to all.
        # Tables WITHOUT contact parameters: the kernels then only flag table contacts (with them they are simulated and a
        # flag is the exception), which gives the regeneration path something to do
        tables_nc = T.tables_from_dict(T.tables_to_dict(tables))
        tables_nc.con_enabled = 0
        gen = SOARM101DataGenerator(_args(tmp), tables=tables_nc, device=rank)
        data = gen.generate_physics_based_data(600, 200, "chirp", seed=5)
        if rank == 0:
            np.save(os.path.join(tmp, "gen.npy"), data)
            np.save(os.path.join(tmp, "gen_flags.npy"), gen.last_flags)
        else:
            assert data is None
        gen.generate_and_save_data()
        assert gen.train_data.shape == (301, 5, 13) and gen.test_data_dict["chirp"].shape == (51, 7, 13)   # on EVERY rank
        # (3) fewer trajectories than ranks: rank 1's shard is empty and nothing hangs
        one = gen.generate_physics_based_data(1, 3, "random", seed=1)
        assert (one.shape == (1, 4, 13)) if rank == 0 else (one is None)
    finally:
        dist.destroy_process_group()


def test_two_rank_dataset_equals_single_process(tmp_path, tables_v):
    _need_two_gpus()
    from lerobot_mujoco_sim2real_b200.SOARM101_DataCollection import SOARM101DataGenerator
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = SOARM101VecEnv(tables=tables_v, num_envs=4097).rollout(7, "chirp", seed=3).cpu().numpy()
    np.testing.assert_array_equal(np.load(tmp_path / "shared.npy"), ref)          # peer-written == one GPU, bit for bit
    d1 = tmp_path / "single"
    from lerobot_mujoco_sim2real_b200 import tables as T_
    tables_nc = T_.tables_from_dict(T_.tables_to_dict(tables_v))
    tables_nc.con_enabled = 0
    gen = SOARM101DataGenerator(_args(d1), tables=tables_nc)
    want = gen.generate_physics_based_data(600, 200, "chirp", seed=5)
    np.testing.assert_array_equal(np.load(tmp_path / "gen.npy"), want)            # incl. the regenerated trajectories
    np.testing.assert_array_equal(np.load(tmp_path / "gen_flags.npy"), gen.last_flags)
    assert gen.last_replaced > 0
    names = sorted(p.name for p in tmp_path.iterdir() if p.name.startswith(("train_", "val_", "test_")))
    assert names == ["test_data_chirp_51_6.npy", "test_data_random_51_6.npy", "test_data_sin_51_6.npy",
                     "train_data_301_4.npy", "val_data_51_6.npy"]
