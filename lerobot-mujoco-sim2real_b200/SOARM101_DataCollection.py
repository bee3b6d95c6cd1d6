"""Drop-in `SOARM101DataGenerator` backed by the fused B200 rollout kernel.

Mirrors [REF SOARM101/SOARM101_DataCollection.py:77-205]: same constructor (`args` with
`xml_path, x_dim, u_dim, device, train_samples, train_steps, test_samples, test_steps,
data_dir_save, data_dir_load_train, data_dir_load_val, env, batch_size, eval_batch_size`), same
`generate_physics_based_data(traj_num, steps, input_type) -> float64 [traj_num, steps+1, 13]`
with columns `[u(5) | ee_pos(3) | qpos(5)]`, same `.npy` cache files and loaders.

Differences, all deliberate:
  * the serial double loop over (trajectory, step) [REF :108-134] is ONE kernel launch per chunk
    of trajectories (`so101_batch_rollout`): reset, control generation, frame_skip x mj_step and
    the row writer are fused; rows stay on the device until the final copy.
  * random numbers come from a counter-based Philox stream keyed (seed, trajectory, step) instead
    of the unseeded global numpy RNG [REF :115,132, train.py:270] — datasets are reproducible and
    independent of batch chunking and of the number of GPUs.
  * under `torch.distributed` the trajectories are sharded across ranks and gathered on rank 0.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

from . import sharding
from . import tables as T
from .vec_env import SOARM101VecEnv


class Collater:
    """[REF SOARM101_DataCollection.py:13-29]"""

    def __init__(self, x_dim: int, u_dim: int, device: str = "cuda"):
        self.x_dim, self.u_dim, self.device = x_dim, u_dim, device

    def __call__(self, batch_list: list):
        batch_data = torch.stack(list(zip(*batch_list))[0], dim=0)
        return dict(x=batch_data[:, :, self.u_dim:self.u_dim + self.x_dim].to(self.device),
                    u=batch_data[:, :, :self.u_dim].to(self.device))


class ContactError(RuntimeError):
    """A generated trajectory left the region the simulator models (see `on_contact`)."""


class SOARM101DataGenerator:
    #: trajectories per launch (bounds device memory: rows are 104 B x (steps+1) per trajectory)
    max_batch = 1 << 20
    #: status bits that make a trajectory unusable as MuJoCo data: state blow-up, or a contact the kernels do not
    #: simulate (SO101_FLAG_TRIP_TABLE is only raised for table contacts the contact path cannot represent)
    BAD_FLAGS = T.FLAG_BADSTATE | T.FLAG_TRIP_TABLE | T.FLAG_TRIP_SELF

    def __init__(self, args, dtype: str = "float64", device: Optional[int] = None, tables=None,
                 gravity_compensation: bool = False, on_contact: str = "regenerate") -> None:
        """gravity_compensation: generate the data as with the (commented-out) gravity-compensation line of
        SOARM101Env.step active [REF SOARM101_Env.py:120] - what the shipped Koopman model was trained on.

        on_contact: what happens to trajectories whose env raised one of `BAD_FLAGS` (the reference scene has contacts
        enabled; a trajectory that reaches an unsimulated contact is NOT what MuJoCo would produce):
          "regenerate" (default)  replace them by fresh trajectories (new env indices, same seed) until none is flagged:
                                  every returned row is valid, the shape stays [traj_num, steps+1, 13];
          "keep"                  return them as they are (`last_flags` holds the mask, a warning is printed, and
                                  `generate_and_save_data` writes the mask next to the dataset as `<file>.flags.npy`);
          "raise"                 raise ContactError."""
        if on_contact not in ("regenerate", "keep", "raise"):
            raise ValueError("on_contact must be 'regenerate', 'keep' or 'raise'")
        self.args = args
        self.gravity_compensation = bool(gravity_compensation)
        self.on_contact = on_contact
        self.udim = self.args.u_dim
        self.xdim = self.args.x_dim
        if self.udim != T.NU_ENV or self.xdim != T.NOBS:
            raise ValueError("SOARM101 data layout is u_dim=5, x_dim=8")
        print("正在初始化物理仿真环境...")
        self.dtype = dtype
        self.device = torch.cuda.current_device() if device is None else device
        if tables is None:
            from .mjcf import attach_tripwire, compile_mjcf
            cm = compile_mjcf(self.args.xml_path)
            attach_tripwire(cm, self.args.xml_path, required=True)
            tables = cm.tables
        self.tables = tables
        self.frame_skip = max(1, int(np.round(0.02 / tables.timestep)))
        self.seed = int(getattr(args, "seed", 42))
        self._calls = 0
        self._envs: Dict[int, SOARM101VecEnv] = {}
        print("物理环境初始化完成。")
        self.collate_fn = Collater(self.args.x_dim, self.args.u_dim, getattr(self.args, "device", "cuda"))
        #: status words of the trajectories of the last call, GLOBAL (all shards) on rank 0, before any replacement
        self.last_flags: Optional[np.ndarray] = None
        #: how many trajectories of the last call were replaced ("regenerate")
        self.last_replaced = 0

    def _env(self, n: int) -> SOARM101VecEnv:
        if n not in self._envs:
            self._envs.clear()
            self._envs[n] = SOARM101VecEnv(tables=self.tables, num_envs=n, dtype=self.dtype, device=self.device)
        return self._envs[n]

    def _rollout_flags(self) -> int:
        return T.ROLL_GRAVCOMP_HOLD if self.gravity_compensation else 0

    def generate_device(self, traj_num: int, steps: int, input_type: str, seed: Optional[int] = None,
                        env_offset: int = 0, out: Optional[torch.Tensor] = None,
                        out_ptr: Optional[int] = None) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        """Trajectories `env_offset .. env_offset + traj_num` as they come out of the kernel ->
        (rows `[traj_num, steps+1, 13]` float64 device tensor, status words `[traj_num]` int32 device tensor).
        out / out_ptr: where the rows go (out_ptr: raw device address, e.g. a slice of `sharding.SharedRows`; the first
        return value is None then).  Nothing is filtered here."""
        if input_type not in ("random", "sin", "chirp"):
            raise ValueError(f"unknown input_type {input_type!r}")
        seed = self.seed if seed is None else seed
        dev = torch.device("cuda", self.device)
        if out_ptr is None and out is None:
            out = torch.empty((traj_num, steps + 1, T.ROW), dtype=torch.float64, device=dev)
        row_bytes = (steps + 1) * T.ROW * 8
        flags = []
        for lo in range(0, traj_num, self.max_batch):
            n = min(self.max_batch, traj_num - lo)
            env = self._env(n)
            if out_ptr is not None:
                env.rollout(steps, input_type, seed=seed, env_offset=env_offset + lo, out_ptr=out_ptr + lo * row_bytes,
                            flags=self._rollout_flags())
            else:
                env.rollout(steps, input_type, seed=seed, env_offset=env_offset + lo, out=out[lo:lo + n],
                            flags=self._rollout_flags())
            flags.append(env.flags())
        fl = torch.cat(flags) if flags else torch.zeros(0, dtype=torch.int32, device=dev)
        return out, fl

    def _replace_flagged(self, rows: torch.Tensor, flags: np.ndarray, steps: int, input_type: str, seed: int) -> int:
        """Overwrite the flagged trajectories of `rows` with flag-free ones simulated from fresh env indices
        (traj_num, traj_num + 1, ... in order, same seed): deterministic, independent of the number of GPUs."""
        bad = np.nonzero(flags & self.BAD_FLAGS)[0]
        total, next_id = len(bad), int(rows.shape[0])
        while len(bad):
            m = len(bad) + len(bad) // 4 + 32                       # oversample: some of the fresh ones trip too
            cand, f = self.generate_device(m, steps, input_type, seed=seed, env_offset=next_id)
            ok = torch.nonzero((f & self.BAD_FLAGS) == 0).flatten()[:len(bad)]
            rows[torch.as_tensor(bad[:len(ok)], device=rows.device)] = cand[ok]
            bad, next_id = bad[len(ok):], next_id + m
        return total

    def generate_physics_based_data(self, traj_num, steps, input_type, seed: Optional[int] = None):
        """[REF SOARM101_DataCollection.py:90-136] -> numpy float64 [traj_num, steps+1, 13].

        With torch.distributed initialised every rank simulates its shard of the trajectories and rank 0 returns the
        full array (other ranks return None): on one node the shards are written straight into rank 0's buffer by the
        rollout kernels (`sharding.SharedRows`), otherwise they are sent point to point (`sharding.gather_rows`).
        Flagged trajectories are handled on rank 0 as `on_contact` says."""
        if seed is None:   # a fresh stream per call, like consecutive draws from the reference's RNG
            seed = self.seed + 7919 * self._calls
        self._calls += 1
        rank, world = sharding.dist_info()
        lo, hi = sharding.shard_range(traj_num, rank, world)
        shared = None
        if world > 1 and sharding.single_node() and torch.distributed.get_backend() == "nccl":
            shared = sharding.SharedRows(traj_num, (steps + 1, T.ROW), torch.float64, self.device, dst=0)
            _, fl = self.generate_device(hi - lo, steps, input_type, seed=seed, env_offset=lo, out_ptr=shared.local_ptr)
            full = shared.finish()
        else:
            local, fl = self.generate_device(hi - lo, steps, input_type, seed=seed, env_offset=lo)
            full = sharding.gather_rows(local, traj_num, dst=0)
        fl_all = sharding.gather_rows(fl, traj_num, dst=0)
        data = None
        try:
            if rank == 0:
                flags = fl_all.cpu().numpy()
                self.last_flags, self.last_replaced = flags, 0
                nbad = int(np.count_nonzero(flags & self.BAD_FLAGS))
                if nbad:
                    what = (f"{nbad} of {traj_num} '{input_type}' trajectories reached a contact the simulator does not "
                            f"model or blew up (table {int(np.count_nonzero(flags & T.FLAG_TRIP_TABLE))}, self "
                            f"{int(np.count_nonzero(flags & T.FLAG_TRIP_SELF))}, bad state "
                            f"{int(np.count_nonzero(flags & T.FLAG_BADSTATE))})")
                    if self.on_contact == "raise":
                        raise ContactError(what)
                    if self.on_contact == "regenerate":
                        self.last_replaced = self._replace_flagged(full, flags, steps, input_type, seed)
                    else:
                        print(f"[so101] WARNING: {what}; kept (on_contact='keep'), mask in last_flags")
                data = full.cpu().numpy()
        finally:
            if shared is not None:
                shared.close()
        return data

    def generate_and_save_data(self):
        """[REF SOARM101_DataCollection.py:138-181]: same cache files, same order.  Under torch.distributed rank 0
        writes the files, every rank then loads them (so all ranks end up with the datasets, like the reference's
        single process)."""
        a = self.args
        os.makedirs(a.data_dir_save, exist_ok=True)
        rank, world = sharding.dist_info()

        def get(path: str, n: int, steps: int, kind: str, what: str):
            exists = os.path.exists(path)
            if world > 1:                      # one decision for all ranks (a half-written file must not split them)
                box = [exists]
                torch.distributed.broadcast_object_list(box, src=0)
                exists = box[0]
            if not exists:
                print(f"生成{what}: {n}条轨迹，每条{steps}步")
                data = self.generate_physics_based_data(n, steps, kind)
                if rank == 0:
                    np.save(path, data)
                    if self.on_contact == "keep" and np.any(self.last_flags & self.BAD_FLAGS):
                        np.save(path[:-4] + ".flags.npy", self.last_flags)
                    print(f"{what}保存到: {path}, 形状: {data.shape}")
                if world > 1:
                    torch.distributed.barrier()        # the file is complete before anybody reads it
                if rank == 0:
                    return data
            return np.load(path)

        self.train_data = get(a.data_dir_load_train, a.train_samples, a.train_steps, "random", "训练数据")
        self.val_data = get(a.data_dir_load_val, a.test_samples, a.test_steps, "random", "验证数据")
        self.test_data_dict = {}
        for test_type in ["random", "sin", "chirp"]:
            path = os.path.join(os.path.dirname(a.data_dir_load_train),
                                f"test_data_{test_type}_{a.test_samples}_{a.test_steps}.npy")
            self.test_data_dict[test_type] = get(path, a.test_samples, a.test_steps, test_type, f"'{test_type}'测试数据")
        self.close()

    def _loader(self, data: np.ndarray, batch_size: int, shuffle: bool) -> DataLoader:
        ds = TensorDataset(torch.tensor(data, dtype=torch.float32))
        return DataLoader(ds, batch_size=batch_size, collate_fn=self.collate_fn, shuffle=shuffle)

    def get_train_loader(self):
        """[REF SOARM101_DataCollection.py:184-196]"""
        return (self._loader(self.train_data, self.args.batch_size, True),
                self._loader(self.val_data, self.args.eval_batch_size, False))

    def get_test_loader(self, test_type):
        """[REF SOARM101_DataCollection.py:198-205]"""
        return self._loader(self.test_data_dict[test_type], self.args.eval_batch_size, False)

    def close(self):
        self._envs.clear()
