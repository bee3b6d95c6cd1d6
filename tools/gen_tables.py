#!/usr/bin/env python
"""Compile the reference MJCF scenes into packed tables committed under the package's assets/.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tools/gen_tables.py [--reference /root/reference]
The JSON files are derived data (about 300 numbers per scene), not reference sources.
tests/test_mjcf.py re-compiles the XML when the reference tree is present and checks that the
committed tables are reproduced bit for bit.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lerobot_mujoco_sim2real_b200 import mjcf, tables  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    so101 = os.path.join(args.reference, "SOARM101", "SO101")
    import numpy as np
    hulls = {}
    for scene, out in tables.BUILTIN_SCENES.items():
        cm = mjcf.compile_mjcf(os.path.join(so101, scene))
        try:
            from lerobot_mujoco_sim2real_b200 import tripwire
            tripwire.fill_tripwire(cm)
            hulls[scene] = tripwire.build_hulls(cm)
        except ImportError:
            pass
        meta = {"source": f"SOARM101/SO101/{scene}", "generator": "tools/gen_tables.py",
                "bodies": cm.body_names, "joints": cm.joint_names, "actuators": cm.actuator_names,
                "sites": cm.site_names, "keys": cm.key_names}
        path = os.path.join(tables.ASSET_DIR, out)
        tables.save_tables(cm.tables, path, meta)
        print("wrote", path)
    # the two scenes carry the same robot: one hull file serves both (checked)
    keys = sorted(hulls)
    for k in keys[1:]:
        for name in hulls[keys[0]]:
            assert np.array_equal(hulls[keys[0]][name], hulls[k][name]), f"hull data differs between scenes: {name}"
    if keys:
        path = os.path.join(tables.ASSET_DIR, tables.BUILTIN_HULLS)
        np.savez_compressed(path, **hulls[keys[0]])
        print("wrote", path, os.path.getsize(path) // 1024, "KiB;", hulls[keys[0]]["vert"].shape[0], "hull vertices")


if __name__ == "__main__":
    main()
