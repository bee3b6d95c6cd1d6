"""The C-ABI library loads on a CPU-only box, exports every symbol include/so101_b200.h declares,
agrees with the Python mirrors on struct layout, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from lerobot_mujoco_sim2real_b200 import _lib, tables as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "so101_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(so101_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in so101_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_header_cites_reference_call_sites():
    src = open(HEADER).read()
    for cite in ("SOARM101_Env.py:34", "SOARM101_Env.py:87-102", "SOARM101_Env.py:128-135",
                 "SOARM101_DataCollection.py:108-134"):
        assert cite in src


def test_struct_layout_and_constants():
    L = _lib.lib()
    assert L.so101_abi_version() == T.ABI_VERSION
    assert L.so101_tables_sizeof() == C.sizeof(T.So101Tables)
    src = open(HEADER).read()
    for name, val in (("SO101_NV", T.NV), ("SO101_MAXBODY", T.MAXBODY), ("SO101_NOBS", T.NOBS),
                      ("SO101_ROW", T.ROW), ("SO101_ABI_VERSION", T.ABI_VERSION), ("SO101_MAXTRIP", T.MAXTRIP)):
        assert re.search(rf"#define {name}\s+{val}\b", src), name
    assert L.so101_batch_state_bytes(1000, T.F64) == 25 * 1000 * 8 + 4000
    assert L.so101_batch_state_bytes(1000, T.F32) == 25 * 1000 * 4 + 4000


def test_model_create_is_host_only_and_validates(tables_v):
    L = _lib.lib()
    h = C.c_void_p()
    assert L.so101_model_create(C.byref(tables_v), C.byref(h)) == 0 and h.value
    L.so101_model_destroy(h)
    bad = T.tables_from_dict(T.tables_to_dict(tables_v))
    bad.body_parent[5] = 2          # break the serial chain
    assert L.so101_model_create(C.byref(bad), C.byref(h)) == -2
    assert b"chain" in L.so101_last_error()
    bad = T.tables_from_dict(T.tables_to_dict(tables_v))
    bad.jnt_solimp[0][4] = 3.0      # unsupported impedance power on a limited joint
    assert L.so101_model_create(C.byref(bad), C.byref(h)) == -2


def test_no_cpu_fallback(tables_v):
    """Without a CUDA device every compute entry point must fail loudly."""
    L = _lib.lib()
    if L.so101_device_count() > 0:
        pytest.skip("a GPU is visible")
    h, b = C.c_void_p(), C.c_void_p()
    assert L.so101_model_create(C.byref(tables_v), C.byref(h)) == 0
    assert L.so101_batch_create(h, 16, T.F64, 0, None, C.byref(b)) == -4
    assert b"no CPU fallback" in L.so101_last_error()
    L.so101_model_destroy(h)
    with pytest.raises(_lib.So101Error):
        from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
        SOARM101VecEnv(tables=tables_v, num_envs=4)
    out = C.c_double()
    assert L.so101_fma_peak(T.F64, 0, C.byref(out)) == -4
    import numpy as np
    A, B, z0 = np.eye(4), np.zeros((4, 2)), np.zeros(4)
    assert L.so101_koopman_score(A.ctypes.data, B.ctypes.data, 4, 2, z0.ctypes.data, None, 1.0, 1.0, None, 0, 1,
                                 T.F64, 0, 0, None, None, None) == -4


def test_product_never_imports_the_oracle():
    """The product path must not import, link, dlopen or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "lerobot-mujoco-sim2real_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(libso101_oracle)|(so101o_)|(oracle\.py)|(oracle/_)", re.M)
    checked = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(src), f"{f} reaches into oracle/"
                checked += 1
    assert checked >= 8
