"""B200-native batched SO-ARM101 stepper (drop-in for SOARM101Env / SOARM101DataGenerator).

Only what the hot path needs lives here: the host MJCF compiler (mjcf.py), the packed tables
(tables.py), the ctypes binding of the CUDA C-ABI library (_lib.py, csrc/), and the host-side
mirrors of the reference interface (SOARM101_Env.py, SOARM101_DataCollection.py, vec_env.py),
plus the device-side twin of the reference's Koopman model / MPC cost (koopman.py), its Cartesian
trajectory generator with the per-way-point inverse kinematics as one launch (TrajectoryGenerator.py, N3) and the
control loop of Koopman_MPC.py over a batch of curves (Koopman_MPC.py).
"""
from . import tables  # noqa: F401
from .tables import builtin_tables, load_tables, save_tables  # noqa: F401

__all__ = ["tables", "builtin_tables", "load_tables", "save_tables"]


def __getattr__(name):  # heavy modules (torch, ctypes library) load on first use
    if name in ("SOARM101VecEnv", "fma_peak_tflops"):
        from . import vec_env
        return getattr(vec_env, name)
    if name == "SOARM101Env":
        from .SOARM101_Env import SOARM101Env
        return SOARM101Env
    if name == "KoopmanModel":
        from .koopman import KoopmanModel
        return KoopmanModel
    if name == "CartesianTrajectoryGenerator":
        from .TrajectoryGenerator import CartesianTrajectoryGenerator
        return CartesianTrajectoryGenerator
    if name == "BatchedKoopmanMPC":
        from .Koopman_MPC import BatchedKoopmanMPC
        return BatchedKoopmanMPC
    if name in ("SOARM101DataGenerator", "Collater"):
        from . import SOARM101_DataCollection as dc
        return getattr(dc, name)
    raise AttributeError(name)
