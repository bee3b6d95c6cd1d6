import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"   # present in the build container only, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REFERENCE)
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def tables_v():
    from lerobot_mujoco_sim2real_b200 import builtin_tables
    return builtin_tables("scene_with_table_v.xml")


@pytest.fixture(scope="session")
def tables_p():
    from lerobot_mujoco_sim2real_b200 import builtin_tables
    return builtin_tables("scene_with_table.xml")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle with the hull data loaded: like the CUDA path (and like MuJoCo on the reference scene) it
    simulates table-plane contact wherever a state reaches the table."""
    from lerobot_mujoco_sim2real_b200 import tables as T
    from oracle import oracle as O
    O.build()
    O.set_hulls(T.builtin_hulls())
    return O


@pytest.fixture()
def contact_free(oracle_mod):
    """For tests of the contact-free pipeline (joint limits, fp32 tolerances, round-1 tripwire semantics): the oracle
    without hull data for the duration of the test; build the CUDA envs with `hulls=None` to match."""
    from lerobot_mujoco_sim2real_b200 import tables as T
    oracle_mod.set_hulls(None)
    yield oracle_mod
    oracle_mod.set_hulls(T.builtin_hulls())
