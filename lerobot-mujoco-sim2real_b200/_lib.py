"""ctypes binding of the CUDA C-ABI library (include/so101_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible,
every compute entry point raises.  The library is built in-tree by build.py
(`python lerobot-mujoco-sim2real_b200/build.py` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from .tables import So101CtrlSpec, So101Hulls, So101IkParams, So101Tables

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libso101_b200.so")

#: every symbol include/so101_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "so101_last_error", "so101_abi_version", "so101_tables_sizeof", "so101_device_count",
    "so101_model_create", "so101_model_destroy", "so101_model_set_hulls",
    "so101_batch_state_bytes", "so101_batch_create", "so101_batch_destroy",
    "so101_batch_reset", "so101_batch_reset_random", "so101_batch_forward",
    "so101_batch_step", "so101_batch_step_flags", "so101_batch_step_host", "so101_batch_reset_host",
    "so101_batch_rollout", "so101_batch_rollout_host", "so101_batch_shoot",
    "so101_batch_get_state", "so101_batch_set_state", "so101_batch_set_qfrc_applied",
    "so101_batch_get_flags", "so101_batch_clear_flags", "so101_batch_stats", "so101_batch_set_option",
    "so101_shared_alloc", "so101_shared_open", "so101_shared_close", "so101_shared_free",
    "so101_fma_peak", "so101_koopman_score", "so101_koopman_create", "so101_koopman_destroy", "so101_koopman_set_gains",
    "so101_koopman_lift", "so101_koopman_feedforward", "so101_koopman_mpc_step", "so101_ik_track",
]


class So101Error(RuntimeError):
    pass


_LIB: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (once) and type the C ABI.  Raises So101Error if the extension is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        # not built yet (fresh checkout): compile it now if nvcc is here; never fall back to a CPU path
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location(
                "_so101_build", os.path.join(os.path.dirname(LIB_PATH), "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        except Exception as exc:
            raise So101Error(
                f"{LIB_PATH} is missing and could not be built ({exc}): run "
                f"`python {os.path.dirname(LIB_PATH)}/build.py` (nvcc, sm_100a).  "
                "There is no CPU fallback for this path.") from exc
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64
    L.so101_last_error.restype = C.c_char_p
    L.so101_abi_version.restype = i32
    L.so101_tables_sizeof.restype = C.c_size_t
    L.so101_device_count.restype = i32
    L.so101_model_create.argtypes = [C.POINTER(So101Tables), C.POINTER(vp)]
    L.so101_model_set_hulls.argtypes = [vp, C.POINTER(So101Hulls)]
    L.so101_model_destroy.argtypes = [vp]
    L.so101_model_destroy.restype = None
    L.so101_batch_state_bytes.argtypes = [i64, i32]
    L.so101_batch_state_bytes.restype = C.c_size_t
    L.so101_batch_create.argtypes = [vp, i64, i32, i32, vp, C.POINTER(vp)]
    L.so101_batch_destroy.argtypes = [vp]
    L.so101_batch_destroy.restype = None
    L.so101_batch_reset.argtypes = [vp, vp, vp, vp, vp]
    L.so101_batch_reset_random.argtypes = [vp, u64, i64, C.c_double, C.c_double, vp, vp]
    L.so101_batch_forward.argtypes = [vp, vp, vp, vp]
    L.so101_batch_step.argtypes = [vp, vp, i32, i32, vp, vp]
    L.so101_batch_step_flags.argtypes = [vp, vp, i32, i32, vp, u32, vp]
    L.so101_batch_step_host.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.so101_batch_reset_host.argtypes = [vp, vp, vp, vp, vp]
    L.so101_batch_rollout.argtypes = [vp, C.POINTER(So101CtrlSpec), i32, i32, vp, u32, vp]
    L.so101_batch_rollout_host.argtypes = [vp, C.POINTER(So101CtrlSpec), vp, i32, i32, vp, u32, vp]
    L.so101_batch_shoot.argtypes = [vp, vp, vp, i32, i32, vp, u32, vp]
    L.so101_batch_get_state.argtypes = [vp, vp, vp, vp, vp]
    L.so101_batch_set_state.argtypes = [vp, vp, vp, vp, vp]
    L.so101_batch_set_qfrc_applied.argtypes = [vp, vp, vp]
    L.so101_batch_get_flags.argtypes = [vp, vp, vp]
    L.so101_batch_clear_flags.argtypes = [vp, vp]
    L.so101_batch_stats.argtypes = [vp, C.POINTER(u64), vp]
    L.so101_batch_set_option.argtypes = [vp, i32, i32]
    L.so101_shared_alloc.argtypes = [i32, C.c_size_t, C.POINTER(vp), C.c_char_p]
    L.so101_shared_open.argtypes = [i32, C.c_char_p, C.POINTER(vp)]
    L.so101_shared_close.argtypes = [i32, vp]
    L.so101_shared_free.argtypes = [i32, vp]
    L.so101_fma_peak.argtypes = [i32, i32, C.POINTER(C.c_double)]
    L.so101_koopman_score.argtypes = [vp, vp, i32, i32, vp, vp, C.c_double, C.c_double, vp, i32, i64, i32, i32, i32, vp, vp, vp]
    L.so101_koopman_create.argtypes = [i32, vp, vp, vp, i32, C.POINTER(vp)]
    L.so101_koopman_destroy.argtypes = [vp]
    L.so101_koopman_destroy.restype = None
    L.so101_koopman_set_gains.argtypes = [vp, i32, i32, vp, vp, vp]
    L.so101_koopman_lift.argtypes = [vp, vp, i32, i32, i64, i64, vp, vp]
    L.so101_koopman_feedforward.argtypes = [vp, vp, i64, i32, vp, vp]
    L.so101_koopman_mpc_step.argtypes = [vp, vp, i32, i32, i64, vp, i64, vp, vp, i32, vp, C.c_double, i64, vp]
    L.so101_ik_track.argtypes = [vp, C.POINTER(So101IkParams), vp, vp, vp, i32, i64, i32, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("so101_abi_version", "so101_device_count"):
            fn.restype = i32
    from .tables import ABI_VERSION
    if L.so101_abi_version() != ABI_VERSION:
        raise So101Error(f"ABI version mismatch: library {L.so101_abi_version()}, python {ABI_VERSION}")
    if L.so101_tables_sizeof() != C.sizeof(So101Tables):
        raise So101Error("So101Tables layout mismatch between tables.py and libso101_b200.so "
                         f"({C.sizeof(So101Tables)} vs {L.so101_tables_sizeof()}): rebuild the library")
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().so101_last_error()
        raise So101Error(f"so101 error {rc}: {msg.decode() if msg else '?'}")


def device_count() -> int:
    return lib().so101_device_count()


def require_device() -> None:
    if device_count() <= 0:
        raise So101Error("no CUDA device visible: the so101 stepper has no CPU fallback")
