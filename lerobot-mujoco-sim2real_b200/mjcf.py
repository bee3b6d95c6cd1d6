"""Host-side MJCF compiler for the SO-ARM101 hinge chain -> packed constant tables.

Replaces `mujoco.MjModel.from_xml_path` [REF SOARM101/SOARM101_Env.py:34] for the subset of
MJCF the hot-path scenes use [REF SOARM101/SO101/scene_with_table_v.xml,
SOARM101/SO101/so101_new_calib_v.xml] (and the position-servo sibling so101_new_calib.xml):

  * <include>, several <compiler>/<option> elements merged in document order
  * nested <default> classes, `class=` / `childclass=` resolution
  * bodies with explicit <inertial> (diaginertia or fullinertia -> principal axes)
  * hinge joints with autolimits, damping / armature / frictionloss / stiffness
  * <general>/<motor>/<position>/<velocity> actuators on joints.  All shortcuts of one default
    class share a single parameter block, so a `<velocity>` without `kv` inherits
    gainprm[0] from a sibling `<position kp=..>` default (kv = 50 here).
  * sites, the first keyframe
  * derived constants MuJoCo computes at compile time: dof_M0, dof_invweight0,
    stat.meaninertia, and `dampratio` -> kv for position-like actuators.

The third-party engine is not vendored in the reference; semantics restated from MuJoCo 3.x
(xml_native_reader.cc OneActuator/OneJoint, user_objects.cc, engine_setconst.c).  Anything
outside the subset raises `MjcfError` instead of being silently ignored when it would change
the dynamics (springs on tendons, equality constraints, non-hinge joints, contacts are
recorded but handled by the tripwire only).
"""
from __future__ import annotations

import copy
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import tables as T


class MjcfError(ValueError):
    pass


# ------------------------------------------------------------------------------------------
# small quaternion toolbox (w, x, y, z), Hamilton product
# ------------------------------------------------------------------------------------------
def q_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def q_norm(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    if n < 1e-15:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def q_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_q(R):
    """Rotation matrix -> unit quaternion (Shepperd)."""
    R = np.asarray(R, dtype=np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return q_norm(q)


def q_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    s = math.sin(0.5 * angle)
    return np.array([math.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def _floats(s: str, n: Optional[int] = None) -> np.ndarray:
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise MjcfError(f"expected {n} numbers, got {s!r}")
    return v


# ------------------------------------------------------------------------------------------
# XML loading: <include> expansion
# ------------------------------------------------------------------------------------------
def _expand_includes(elem: ET.Element, base_dir: str, depth: int = 0) -> None:
    if depth > 16:
        raise MjcfError("include nesting too deep")
    i = 0
    while i < len(elem):
        ch = elem[i]
        if ch.tag == "include":
            path = os.path.join(base_dir, ch.attrib["file"])
            sub = ET.parse(path).getroot()
            _expand_includes(sub, os.path.dirname(path), depth + 1)
            elem.remove(ch)
            for k, g in enumerate(list(sub)):
                elem.insert(i + k, g)
            i += len(sub)
        else:
            _expand_includes(ch, base_dir, depth)
            i += 1


def load_xml(path: str) -> ET.Element:
    if not os.path.exists(path):
        raise FileNotFoundError(f"MuJoCo XML 文件未找到: {path}")  # same text as the reference Env
    root = ET.parse(path).getroot()
    if root.tag != "mujoco":
        raise MjcfError("root element must be <mujoco>")
    _expand_includes(root, os.path.dirname(os.path.abspath(path)))
    return root


# ------------------------------------------------------------------------------------------
# defaults
# ------------------------------------------------------------------------------------------
@dataclass
class ActuatorParams:
    """The single actuator parameter block a default class owns (shared by all shortcuts)."""
    gainprm: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0]))
    biasprm: np.ndarray = field(default_factory=lambda: np.zeros(3))
    gaintype: str = "fixed"
    biastype: str = "none"
    dyntype: str = "none"
    ctrllimited: str = "auto"
    forcelimited: str = "auto"
    ctrlrange: np.ndarray = field(default_factory=lambda: np.zeros(2))
    forcerange: np.ndarray = field(default_factory=lambda: np.zeros(2))
    gear: float = 1.0


_ACT_TAGS = ("general", "motor", "position", "velocity")


def _apply_actuator(p: ActuatorParams, tag: str, a: Dict[str, str]) -> None:
    """One actuator element (in <default> or <actuator>) applied onto a parameter block.

    Restates xml_native_reader.cc OneActuator + mjs_setTo{Motor,Position,Velocity}.
    """
    if "ctrllimited" in a:
        p.ctrllimited = a["ctrllimited"]
    if "forcelimited" in a:
        p.forcelimited = a["forcelimited"]
    if "ctrlrange" in a:
        p.ctrlrange = _floats(a["ctrlrange"], 2)
    if "forcerange" in a:
        p.forcerange = _floats(a["forcerange"], 2)
    if "gear" in a:
        p.gear = float(_floats(a["gear"])[0])
    for bad in ("actlimited", "actrange", "timeconst", "inheritrange", "actdim", "dynprm"):
        if bad in a:
            raise MjcfError(f"actuator attribute {bad!r} (activation dynamics) is not supported")
    if tag == "general":
        if "gainprm" in a:
            g = _floats(a["gainprm"])
            p.gainprm[: min(3, g.size)] = g[:3]
        if "biasprm" in a:
            b = _floats(a["biasprm"])
            p.biasprm[: min(3, b.size)] = b[:3]
        p.gaintype = a.get("gaintype", p.gaintype)
        p.biastype = a.get("biastype", p.biastype)
        p.dyntype = a.get("dyntype", p.dyntype)
    elif tag == "motor":
        p.gainprm[:] = (1.0, 0, 0)
        p.biasprm[:] = 0
        p.gaintype, p.biastype, p.dyntype = "fixed", "none", "none"
    elif tag == "position":
        kp = float(a["kp"]) if "kp" in a else p.gainprm[0]
        if "kv" in a and "dampratio" in a:
            raise MjcfError("kv and dampratio cannot both be defined")
        p.gainprm[0] = kp
        p.biasprm[0] = 0.0
        p.biasprm[1] = -kp
        if "kv" in a:
            if float(a["kv"]) < 0:
                raise MjcfError("kv cannot be negative")
            p.biasprm[2] = -float(a["kv"])
        if "dampratio" in a:
            if float(a["dampratio"]) < 0:
                raise MjcfError("dampratio cannot be negative")
            p.biasprm[2] = float(a["dampratio"])  # positive = damping ratio, resolved below
        p.gaintype, p.biastype, p.dyntype = "fixed", "affine", "none"
    elif tag == "velocity":
        kv = float(a["kv"]) if "kv" in a else p.gainprm[0]   # <- F2: inherits the class's kp
        p.biasprm[:] = 0
        p.gainprm[0] = kv
        p.biasprm[2] = -kv
        p.gaintype, p.biastype, p.dyntype = "fixed", "affine", "none"
    else:
        raise MjcfError(f"actuator type <{tag}> is not supported")


@dataclass
class DefaultClass:
    name: str
    joint: Dict[str, str] = field(default_factory=dict)
    geom: Dict[str, str] = field(default_factory=dict)
    site: Dict[str, str] = field(default_factory=dict)
    actuator: ActuatorParams = field(default_factory=ActuatorParams)


def _read_defaults(root: ET.Element) -> Dict[str, DefaultClass]:
    classes: Dict[str, DefaultClass] = {"main": DefaultClass("main")}

    def visit(elem: ET.Element, parent: Optional[DefaultClass], top: bool) -> None:
        name = elem.attrib.get("class")
        if top:
            cls = classes["main"]
            if name not in (None, "main"):
                raise MjcfError("top-level default class must be unnamed or 'main'")
        else:
            if not name:
                raise MjcfError("nested <default> needs a class name")
            if name in classes:
                raise MjcfError(f"repeated default class {name!r}")
            cls = copy.deepcopy(parent)
            cls.name = name
            classes[name] = cls
        for ch in elem:
            if ch.tag == "default":
                continue
            if ch.tag == "joint":
                cls.joint.update(ch.attrib)
            elif ch.tag == "geom":
                cls.geom.update(ch.attrib)
            elif ch.tag == "site":
                cls.site.update(ch.attrib)
            elif ch.tag in _ACT_TAGS:
                _apply_actuator(cls.actuator, ch.tag, ch.attrib)
            elif ch.tag in ("mesh", "material", "camera", "light"):
                pass  # visual only
            else:
                raise MjcfError(f"<default><{ch.tag}> is not supported")
        for ch in elem:
            if ch.tag == "default":
                visit(ch, cls, False)

    for d in root.findall("default"):
        visit(d, None, True)
    return classes


# ------------------------------------------------------------------------------------------
# compiled intermediate model
# ------------------------------------------------------------------------------------------
@dataclass
class Geom:
    name: str
    body: int
    type: str
    size: np.ndarray
    pos: np.ndarray
    quat: np.ndarray
    mesh: Optional[str]
    contype: int
    conaffinity: int
    # contact parameters (MuJoCo defaults: friction "1 0.005 0.0001", solref "0.02 1", solimp "0.9 0.95 0.001 0.5 2")
    friction: np.ndarray = None
    solref: np.ndarray = None
    solimp: np.ndarray = None
    margin: float = 0.0
    gap: float = 0.0
    condim: int = 3
    priority: int = 0
    solmix: float = 1.0


@dataclass
class CompiledModel:
    """Tables plus the name maps the Env shims need (joint(name).id, site ids, geoms)."""
    tables: T.So101Tables
    body_names: List[str]
    joint_names: List[str]
    actuator_names: List[str]
    site_names: List[str]
    site_body: List[int]
    site_pos: List[np.ndarray]
    key_names: List[str]
    geoms: List[Geom]
    mesh_files: Dict[str, str]
    meshdir: str
    xml_dir: str
    M0: np.ndarray
    hulls: Optional[dict] = None     # convex hulls of the colliding geoms (tripwire.build_hulls); None: the built-in ones

    def joint_id(self, name: str) -> int:
        return self.joint_names.index(name)

    def site_id(self, name: str) -> int:
        return self.site_names.index(name)


def _frame_quat(a: Dict[str, str], angle_scale: float, eulerseq: str) -> np.ndarray:
    if "quat" in a:
        return q_norm(_floats(a["quat"], 4))
    if "euler" in a:
        e = _floats(a["euler"], 3) * angle_scale
        q = np.array([1.0, 0, 0, 0])
        for k, ax in enumerate(eulerseq):
            axis = {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}[ax.lower()]
            qa = q_axis_angle(axis, e[k])
            q = q_mul(q, qa) if ax.islower() else q_mul(qa, q)
        return q_norm(q)
    if "axisangle" in a:
        v = _floats(a["axisangle"], 4)
        ax = v[:3] / np.linalg.norm(v[:3])
        return q_axis_angle(ax, v[3] * angle_scale)
    for bad in ("xyaxes", "zaxis"):
        if bad in a:
            raise MjcfError(f"orientation attribute {bad!r} is not supported")
    return np.array([1.0, 0, 0, 0])


def _principal(full6: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """fullinertia (Ixx Iyy Izz Ixy Ixz Iyz) -> principal moments (descending) + frame quat."""
    Ixx, Iyy, Izz, Ixy, Ixz, Iyz = full6
    I = np.array([[Ixx, Ixy, Ixz], [Ixy, Iyy, Iyz], [Ixz, Iyz, Izz]])
    w, V = np.linalg.eigh(I)
    order = np.argsort(-w)
    w, V = w[order], V[:, order]
    if np.linalg.det(V) < 0:
        V[:, 2] = -V[:, 2]
    if w[2] <= 0:
        raise MjcfError("inertia must be positive definite")
    if w[0] > w[1] + w[2] + 1e-12 * w[0]:
        raise MjcfError("inertia must satisfy A + B >= C")
    return w, mat_q(V)


_TRIP_FIELDS = ("ntrip", "nself", "trip_body", "trip_center", "trip_axes", "trip_half", "trip_plane_z", "trip_qbox",
                "con_friction", "con_solref", "con_solimp", "con_margin", "con_box", "con_enabled", "con_condim")


def attach_tripwire(cm: "CompiledModel", xml_path: str, required: bool = False) -> str:
    """Fill the contact-tripwire tables of a freshly compiled model.

    Fast path: the scene is one of the built-in ones and compiles to the committed tables bit for bit
    -> reuse their precomputed tripwire.  Otherwise build it from the collision meshes (tripwire.py,
    needs scipy and the STL files, ~15 s).  Returns how it was obtained; "none" leaves the tripwire
    disabled (no env is ever flagged) and says why on stderr; with `required` (the data generator) that is an error:
    the scene has contacts enabled and trajectories that reach one must be recognisable."""
    import ctypes
    import sys
    key = os.path.basename(xml_path)
    if key in T.BUILTIN_SCENES:
        try:
            ref = T.builtin_tables(key)
            probe = T.So101Tables.from_buffer_copy(bytes(cm.tables))
            for f in _TRIP_FIELDS:
                setattr(probe, f, getattr(ref, f))
            if bytes(probe) == bytes(ref):
                ctypes.memmove(ctypes.byref(cm.tables), ctypes.byref(ref), ctypes.sizeof(ref))
                return "builtin"
        except Exception:
            pass
    try:
        from . import tripwire
        tripwire.fill_tripwire(cm)
        cm.hulls = tripwire.build_hulls(cm)
        return "meshes"
    except Exception as exc:   # scipy or meshes missing: stepping still works, contact is just not flagged
        if required:
            raise MjcfError(f"contact tables for {xml_path} could not be built ({exc}); without them trajectories that "
                            "touch the table would be written to the dataset unnoticed") from exc
        print(f"[so101] contact tripwire disabled for {xml_path}: {exc}", file=sys.stderr)
        return "none"


def compile_mjcf(xml_path: str, site_name: str = "gripperframe") -> CompiledModel:
    root = load_xml(xml_path)
    xml_dir = os.path.dirname(os.path.abspath(xml_path))

    # ---- <compiler>: merged in document order ---------------------------------------------
    comp = {"angle": "degree", "autolimits": "true", "eulerseq": "xyz", "meshdir": "",
            "inertiafromgeom": "auto", "coordinate": "local"}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    if comp["coordinate"] != "local":
        raise MjcfError("global coordinates are not supported")
    if comp["inertiafromgeom"] == "true":
        raise MjcfError("inertiafromgeom=true is not supported (explicit <inertial> required)")
    angle_scale = 1.0 if comp["angle"] == "radian" else math.pi / 180.0
    autolimits = comp["autolimits"] == "true"

    # ---- <option> --------------------------------------------------------------------------
    opt = {"timestep": "0.002", "gravity": "0 0 -9.81", "tolerance": "1e-8", "ls_tolerance": "0.01",
           "iterations": "100", "ls_iterations": "50", "integrator": "Euler", "solver": "Newton",
           "cone": "pyramidal", "impratio": "1", "jacobian": "auto", "noslip_iterations": "0"}
    for o in root.findall("option"):
        opt.update(o.attrib)
        for fl in o.findall("flag"):
            for k, v in fl.attrib.items():
                dflt = "disable" if k in ("override", "energy", "fwdinv", "invdiscrete", "multiccd",
                                           "island", "nativeccd") else "enable"
                if v != dflt:
                    raise MjcfError(f"<flag {k}={v!r}> changes the pipeline and is not supported")
    if opt["integrator"] != "Euler":
        raise MjcfError(f"integrator {opt['integrator']!r} is not supported (scene uses Euler)")
    if opt["solver"] != "Newton":
        raise MjcfError(f"solver {opt['solver']!r} is not supported (scene uses Newton)")
    if int(opt["noslip_iterations"]) != 0:
        raise MjcfError("noslip iterations are not supported")
    for sect in ("equality", "tendon"):
        for e in root.findall(sect):
            if len(e):
                raise MjcfError(f"<{sect}> elements are not supported")

    classes = _read_defaults(root)

    def cls_of(elem: ET.Element, childclass: Optional[str]) -> DefaultClass:
        name = elem.attrib.get("class") or childclass or "main"
        if name not in classes:
            raise MjcfError(f"unknown default class {name!r}")
        return classes[name]

    # ---- assets (mesh file names, for the tripwire) ---------------------------------------------
    mesh_files: Dict[str, str] = {}
    for asset in root.findall("asset"):
        for m in asset.findall("mesh"):
            f = m.attrib.get("file")
            if f:
                mesh_files[m.attrib.get("name") or os.path.splitext(os.path.basename(f))[0]] = f

    # ---- kinematic tree, depth first in document order ------------------------------------------
    body_names = ["world"]
    body_parent = [0]
    body_pos = [np.zeros(3)]
    body_quat = [np.array([1.0, 0, 0, 0])]
    body_ipos = [np.zeros(3)]
    body_iquat = [np.array([1.0, 0, 0, 0])]
    body_inertia = [np.zeros(3)]
    body_mass = [0.0]
    body_jnt = [-1]
    joints: List[dict] = []
    sites: List[Tuple[str, int, np.ndarray, np.ndarray]] = []
    geoms: List[Geom] = []

    def read_geoms_sites(elem: ET.Element, bid: int, childclass: Optional[str]) -> None:
        for g in elem.findall("geom"):
            a = dict(cls_of(g, childclass).geom)
            a.update(g.attrib)
            geoms.append(Geom(
                name=a.get("name", ""), body=bid, type=a.get("type", "sphere"),
                size=_floats(a.get("size", "0 0 0")),
                pos=_floats(a.get("pos", "0 0 0"), 3), quat=_frame_quat(a, angle_scale, comp["eulerseq"]),
                mesh=a.get("mesh"), contype=int(a.get("contype", "1")),
                conaffinity=int(a.get("conaffinity", "1")),
                friction=_pad(_floats(a.get("friction", "1 0.005 0.0001")), [1.0, 0.005, 0.0001]),
                solref=_pad(_floats(a.get("solref", "0.02 1")), [0.02, 1.0]),
                solimp=_pad_solimp(_floats(a.get("solimp", "0.9 0.95 0.001 0.5 2"))),
                margin=float(a.get("margin", "0")), gap=float(a.get("gap", "0")), condim=int(a.get("condim", "3")),
                priority=int(a.get("priority", "0")), solmix=float(a.get("solmix", "1"))))
        for s in elem.findall("site"):
            a = dict(cls_of(s, childclass).site)
            a.update(s.attrib)
            sites.append((a.get("name", ""), bid, _floats(a.get("pos", "0 0 0"), 3),
                          _frame_quat(a, angle_scale, comp["eulerseq"])))

    def visit_body(elem: ET.Element, parent: int, childclass: Optional[str]) -> None:
        a = elem.attrib
        childclass = a.get("childclass", childclass)
        if a.get("mocap", "false") == "true":
            raise MjcfError("mocap bodies are not supported")
        bid = len(body_names)
        body_names.append(a.get("name", f"body{bid}"))
        body_parent.append(parent)
        body_pos.append(_floats(a.get("pos", "0 0 0"), 3))
        body_quat.append(_frame_quat(a, angle_scale, comp["eulerseq"]))
        inert = elem.find("inertial")
        if inert is None:
            raise MjcfError(f"body {body_names[-1]!r} has no <inertial> (inertia from geoms unsupported)")
        ia = inert.attrib
        mass = float(ia["mass"])
        iq = _frame_quat(ia, angle_scale, comp["eulerseq"])
        if "fullinertia" in ia:
            if "quat" in ia or "euler" in ia:
                raise MjcfError("fullinertia with an orientation is not supported")
            mom, iq = _principal(_floats(ia["fullinertia"], 6))
        elif "diaginertia" in ia:
            mom = _floats(ia["diaginertia"], 3)
        else:
            raise MjcfError("<inertial> needs diaginertia or fullinertia")
        body_ipos.append(_floats(ia.get("pos", "0 0 0"), 3))
        body_iquat.append(iq)
        body_inertia.append(mom)
        body_mass.append(mass)
        if elem.find("freejoint") is not None:
            raise MjcfError("free joints are not supported")
        jl = elem.findall("joint")
        if len(jl) > 1:
            raise MjcfError("more than one joint per body is not supported")
        body_jnt.append(-1)
        if jl:
            j = jl[0]
            ja = dict(cls_of(j, childclass).joint)
            ja.update(j.attrib)
            if ja.get("type", "hinge") != "hinge":
                raise MjcfError(f"joint type {ja.get('type')!r} is not supported (hinge chain only)")
            axis = _floats(ja.get("axis", "0 0 1"), 3)
            axis = axis / np.linalg.norm(axis)
            has_range = "range" in ja
            rng = _floats(ja["range"], 2) * angle_scale if has_range else np.zeros(2)
            lim = ja.get("limited", "auto")
            if lim == "auto":
                if has_range and not autolimits:
                    raise MjcfError("joint has range but limited is unspecified and autolimits is off")
                limited = has_range
            else:
                limited = lim == "true"
            body_jnt[bid] = len(joints)
            joints.append(dict(
                name=ja.get("name", f"joint{len(joints)}"), body=bid, axis=axis,
                pos=_floats(ja.get("pos", "0 0 0"), 3), range=rng, limited=int(limited),
                margin=float(ja.get("margin", "0")),
                solreflimit=_floats(ja.get("solreflimit", "0.02 1"), 2),
                solimplimit=_pad_solimp(_floats(ja.get("solimplimit", "0.9 0.95 0.001 0.5 2"))),
                solreffriction=_floats(ja.get("solreffriction", "0.02 1"), 2),
                solimpfriction=_pad_solimp(_floats(ja.get("solimpfriction", "0.9 0.95 0.001 0.5 2"))),
                stiffness=float(ja.get("stiffness", "0")),
                ref=float(ja.get("ref", "0")) * angle_scale,
                springref=float(ja.get("springref", "0")) * angle_scale,
                armature=float(ja.get("armature", "0")), damping=float(ja.get("damping", "0")),
                frictionloss=float(ja.get("frictionloss", "0"))))
            if "springdamper" in ja:
                raise MjcfError("springdamper is not supported")
        if float(a.get("gravcomp", "0")) != 0:
            raise MjcfError("body gravcomp is not supported")
        read_geoms_sites(elem, bid, childclass)
        for ch in elem.findall("body"):
            visit_body(ch, bid, childclass)

    for wb in root.findall("worldbody"):
        read_geoms_sites(wb, 0, None)
        for b in wb.findall("body"):
            visit_body(b, 0, None)

    nbody, nv = len(body_names), len(joints)
    if nv != T.NV:
        raise MjcfError(f"this build is specialised for {T.NV} hinge dofs, the model has {nv}")
    if nbody > T.MAXBODY:
        raise MjcfError(f"too many bodies ({nbody} > {T.MAXBODY})")

    # ---- actuators ------------------------------------------------------------------------------
    acts: List[dict] = []
    for sect in root.findall("actuator"):
        for e in sect:
            if e.tag not in _ACT_TAGS:
                raise MjcfError(f"actuator type <{e.tag}> is not supported")
            p = copy.deepcopy(cls_of(e, None).actuator)
            _apply_actuator(p, e.tag, e.attrib)
            if "joint" not in e.attrib:
                raise MjcfError("only joint transmissions are supported")
            jn = e.attrib["joint"]
            names = [j["name"] for j in joints]
            if jn not in names:
                raise MjcfError(f"actuator refers to unknown joint {jn!r}")
            if p.gaintype != "fixed" or p.biastype not in ("none", "affine") or p.dyntype != "none":
                raise MjcfError("only fixed-gain / affine-bias / stateless actuators are supported")

            def lim(flag: str, rng: np.ndarray, what: str) -> int:
                if flag == "auto":
                    has = bool(rng[0] != 0 or rng[1] != 0)
                    if has and not autolimits:
                        raise MjcfError(f"{what} given but autolimits is off")
                    return int(has)
                return int(flag == "true")

            acts.append(dict(
                name=e.attrib.get("name", f"actuator{len(acts)}"), dof=names.index(jn), gear=p.gear,
                gain=float(p.gainprm[0]),
                bias=(p.biasprm.copy() if p.biastype == "affine" else np.zeros(3)),
                ctrlrange=p.ctrlrange.copy(), forcerange=p.forcerange.copy(),
                ctrllimited=lim(p.ctrllimited, p.ctrlrange, "ctrlrange"),
                forcelimited=lim(p.forcelimited, p.forcerange, "forcerange")))
    if len(acts) != T.NV:
        raise MjcfError(f"expected {T.NV} actuators, found {len(acts)}")
    if sorted(a["dof"] for a in acts) != list(range(T.NV)):
        raise MjcfError("expected exactly one actuator per joint")

    # ---- tables ----------------------------------------------------------------------------------
    t = T.So101Tables()
    t.abi_version = T.ABI_VERSION
    t.nbody, t.nv, t.nu = nbody, nv, len(acts)
    t.iterations, t.ls_iterations = int(opt["iterations"]), int(opt["ls_iterations"])
    t.timestep = float(opt["timestep"])
    t.gravity[:] = list(_floats(opt["gravity"], 3))
    t.tolerance, t.ls_tolerance = float(opt["tolerance"]), float(opt["ls_tolerance"])
    for b in range(nbody):
        t.body_parent[b] = body_parent[b]
        t.body_jnt[b] = body_jnt[b]
        t.body_pos[b][:] = list(body_pos[b])
        t.body_quat[b][:] = list(body_quat[b])
        t.body_ipos[b][:] = list(body_ipos[b])
        t.body_iquat[b][:] = list(body_iquat[b])
        t.body_inertia[b][:] = list(body_inertia[b])
        t.body_mass[b] = body_mass[b]
    for b in range(nbody, T.MAXBODY):
        t.body_jnt[b] = -1
        t.body_quat[b][0] = 1.0
        t.body_iquat[b][0] = 1.0
    for k, j in enumerate(joints):
        t.jnt_body[k] = j["body"]
        t.jnt_limited[k] = j["limited"]
        t.jnt_pos[k][:] = list(j["pos"])
        t.jnt_axis[k][:] = list(j["axis"])
        t.jnt_range[k][:] = list(j["range"])
        t.jnt_margin[k] = j["margin"]
        t.jnt_solref[k][:] = list(j["solreflimit"])
        t.jnt_solimp[k][:] = list(j["solimplimit"])
        t.jnt_stiffness[k] = j["stiffness"]
        t.qpos0[k] = j["ref"]
        t.qpos_spring[k] = j["springref"]
        t.dof_armature[k] = j["armature"]
        t.dof_damping[k] = j["damping"]
        t.dof_frictionloss[k] = j["frictionloss"]
        t.dof_solref[k][:] = list(j["solreffriction"])
        t.dof_solimp[k][:] = list(j["solimpfriction"])
    # joint k must be a descendant chain for the CUDA path; checked again by so101_model_create
    by_dof = sorted(acts, key=lambda a: a["dof"])
    actuator_names = [a["name"] for a in acts]
    for k, a in enumerate(acts):
        t.act_dof[k] = a["dof"]
        t.act_ctrllimited[k] = a["ctrllimited"]
        t.act_forcelimited[k] = a["forcelimited"]
        t.act_gear[k] = a["gear"]
        t.act_gain[k] = a["gain"]
        t.act_bias[k][:] = list(a["bias"])
        t.act_ctrlrange[k][:] = list(a["ctrlrange"])
        t.act_forcerange[k][:] = list(a["forcerange"])
    del by_dof

    site_names = [s[0] for s in sites]
    if site_name not in site_names:
        raise MjcfError("在模型中未找到名为 'gripper' 的 site。请检查XML文件。")  # reference text
    sid = site_names.index(site_name)
    t.site_body = sites[sid][1]
    t.site_pos[:] = list(sites[sid][2])
    t.site_quat[:] = list(sites[sid][3])

    # ---- derived constants (engine_setconst.c): M(qpos0), invweight0, meaninertia, dampratio ------
    M0 = mass_matrix_numpy(t, np.array([j["ref"] for j in joints]))
    Minv = np.linalg.inv(M0)
    for k in range(nv):
        t.dof_M0[k] = M0[k, k]
        t.dof_invweight0[k] = Minv[k, k]
    t.meaninertia = float(np.mean(np.diag(M0)))
    # body_invweight0 (engine_setconst.c set0): mean translational / rotational diagonal of J M^-1 J' with J the
    # Jacobian of the body's centre of mass at qpos0; bodies without dofs above them get 0
    q0 = np.array([j["ref"] for j in joints])
    xpos0, xmat0, anchor0, axis0 = fk_numpy(t, q0)
    for b in range(1, nbody):
        com = xpos0[b] + xmat0[b] @ np.array(t.body_ipos[b][:])
        Jp, Jr = np.zeros((3, nv)), np.zeros((3, nv))
        a = b
        while a > 0:
            k = t.body_jnt[a]
            if k >= 0:
                Jp[:, k] = np.cross(axis0[k], com - anchor0[k])
                Jr[:, k] = axis0[k]
            a = t.body_parent[a]
        t.body_invweight0[b][0] = float(np.trace(Jp @ Minv @ Jp.T) / 3.0)
        t.body_invweight0[b][1] = float(np.trace(Jr @ Minv @ Jr.T) / 3.0)
    for k, a in enumerate(acts):
        gain, b = t.act_gain[k], t.act_bias[k]
        if gain == -b[1] and b[2] > 0:   # position-like actuator carrying a damping ratio
            mass = t.dof_M0[a["dof"]] / (a["gear"] * a["gear"])
            b[2] = -(b[2] * 2.0 * math.sqrt(gain * mass))

    # ---- keyframe 0 --------------------------------------------------------------------------------
    key_names: List[str] = []
    for ks in root.findall("keyframe"):
        for kf in ks.findall("key"):
            if not key_names:
                q = _floats(kf.attrib["qpos"], nv) if "qpos" in kf.attrib else np.array(t.qpos0[:])
                c = _floats(kf.attrib["ctrl"], nv) if "ctrl" in kf.attrib else np.zeros(nv)
                t.key_qpos[:] = list(q)
                t.key_ctrl[:] = list(c)
            key_names.append(kf.attrib.get("name", f"key{len(key_names)}"))
    if not key_names:
        t.key_qpos[:] = list(t.qpos0[:])

    # tripwire defaults: disabled until tools/gen_tables.py fills it from the collision meshes
    t.ntrip = 0
    t.nself = 0
    t.trip_plane_z = -1e30
    for k in range(nv):
        t.trip_qbox[k][0], t.trip_qbox[k][1] = -1e30, 1e30

    return CompiledModel(
        tables=t, body_names=body_names, joint_names=[j["name"] for j in joints],
        actuator_names=actuator_names, site_names=site_names, site_body=[s[1] for s in sites],
        site_pos=[s[2] for s in sites], key_names=key_names, geoms=geoms, mesh_files=mesh_files,
        meshdir=os.path.join(xml_dir, comp.get("meshdir", "")), xml_dir=xml_dir, M0=M0)


def _pad(v: np.ndarray, full) -> np.ndarray:
    out = np.array(full, dtype=np.float64)
    out[: v.size] = v
    return out


def _pad_solimp(v: np.ndarray) -> np.ndarray:
    full = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
    full[: v.size] = v
    return full


# ------------------------------------------------------------------------------------------
# numpy kinematics / mass matrix (compile-time constants; also an independent check in tests)
# ------------------------------------------------------------------------------------------
def fk_numpy(t: T.So101Tables, q: np.ndarray):
    """World frames of every body at joint vector q.  Returns (xpos, xmat, anchor, axis)."""
    nb = t.nbody
    xpos = np.zeros((nb, 3))
    xmat = np.tile(np.eye(3), (nb, 1, 1))
    anchor = np.zeros((t.nv, 3))
    axis = np.zeros((t.nv, 3))
    for b in range(1, nb):
        p = t.body_parent[b]
        pos = xpos[p] + xmat[p] @ np.array(t.body_pos[b][:])
        R = xmat[p] @ q_mat(q_norm(np.array(t.body_quat[b][:])))
        j = t.body_jnt[b]
        if j >= 0:
            jp = np.array(t.jnt_pos[j][:])
            ja = np.array(t.jnt_axis[j][:])
            anchor[j] = pos + R @ jp
            axis[j] = R @ ja
            Rj = q_mat(q_axis_angle(ja, q[j] - t.qpos0[j]))
            R = R @ Rj
            pos = anchor[j] - R @ jp
        xpos[b], xmat[b] = pos, R
    return xpos, xmat, anchor, axis


def mass_matrix_numpy(t: T.So101Tables, q: np.ndarray) -> np.ndarray:
    """M(q) = sum_b m Jv^T Jv + Jw^T I_world Jw + diag(armature) (textbook form)."""
    xpos, xmat, anchor, axis = fk_numpy(t, q)
    nv = t.nv
    M = np.diag(np.array(t.dof_armature[:]))
    for b in range(1, t.nbody):
        m = t.body_mass[b]
        com = xpos[b] + xmat[b] @ np.array(t.body_ipos[b][:])
        Ri = xmat[b] @ q_mat(q_norm(np.array(t.body_iquat[b][:])))
        Iw = Ri @ np.diag(np.array(t.body_inertia[b][:])) @ Ri.T
        Jv = np.zeros((3, nv))
        Jw = np.zeros((3, nv))
        a = b
        while a > 0:
            j = t.body_jnt[a]
            if j >= 0:
                Jw[:, j] = axis[j]
                Jv[:, j] = np.cross(axis[j], com - anchor[j])
            a = t.body_parent[a]
        M = M + m * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


def gravity_bias_numpy(t: T.So101Tables, q: np.ndarray) -> np.ndarray:
    """qfrc_bias at zero velocity = -J_com^T m g summed over bodies."""
    xpos, xmat, anchor, axis = fk_numpy(t, q)
    g = np.array(t.gravity[:])
    out = np.zeros(t.nv)
    for b in range(1, t.nbody):
        com = xpos[b] + xmat[b] @ np.array(t.body_ipos[b][:])
        a = b
        while a > 0:
            j = t.body_jnt[a]
            if j >= 0:
                out[j] -= t.body_mass[b] * np.dot(np.cross(axis[j], com - anchor[j]), g)
            a = t.body_parent[a]
    return out


def site_numpy(t: T.So101Tables, q: np.ndarray) -> np.ndarray:
    xpos, xmat, _, _ = fk_numpy(t, q)
    return xpos[t.site_body] + xmat[t.site_body] @ np.array(t.site_pos[:])
