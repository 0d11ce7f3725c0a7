"""ctypes binding of the C ABI in include/cpb200.h (libcpb200.so, the CUDA step engine).

This is plumbing for the tests and the benchmark: it converts scene blobs
(scenes/cpb_scene.h) into the flat descriptors the C ABI takes and moves numpy
buffers in and out.  The product's host side is the C99 layer in host/ (the
Chipmunk2D public API); both sit on the same C ABI.

There is no CPU path: constructing a World without a usable CUDA device raises.
"""
import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "lib")
ENGINE_LIB = os.path.join(LIB_DIR, "libcpb200.so")

# ---- numpy mirrors of the C structs (align=True reproduces the C layout) -----------

SCENE_HEADER = np.dtype([
    ("magic", "<u8"), ("n_bodies", "<i4"), ("n_shapes", "<i4"), ("n_verts", "<i4"), ("n_joints", "<i4"),
    ("iterations", "<i4"), ("collision_persistence", "<u4"), ("gravity", "<f8", 2), ("damping", "<f8"),
    ("idle_speed_threshold", "<f8"), ("sleep_time_threshold", "<f8"), ("collision_slop", "<f8"),
    ("collision_bias", "<f8"), ("timestep", "<f8")], align=True)
SCENE_BODY = np.dtype([
    ("type", "<i4"), ("is_space_static", "<i4"), ("m", "<f8"), ("i", "<f8"), ("cog", "<f8", 2),
    ("p", "<f8", 2), ("v", "<f8", 2), ("f", "<f8", 2), ("a", "<f8"), ("w", "<f8"), ("t", "<f8")], align=True)
SCENE_SHAPE = np.dtype([
    ("type", "<i4"), ("body", "<i4"), ("sensor", "<i4"), ("n_verts", "<i4"), ("vert_offset", "<i4"),
    ("categories", "<u4"), ("mask", "<u4"), ("_pad", "<i4"), ("group", "<u8"), ("collision_type", "<u8"),
    ("e", "<f8"), ("u", "<f8"), ("surface_v", "<f8", 2), ("r", "<f8"), ("a", "<f8", 2), ("b", "<f8", 2),
    ("a_tangent", "<f8", 2), ("b_tangent", "<f8", 2), ("mass", "<f8")], align=True)
SCENE_JOINT = np.dtype([
    ("type", "<i4"), ("a", "<i4"), ("b", "<i4"), ("collide_bodies", "<i4"), ("max_force", "<f8"),
    ("error_bias", "<f8"), ("max_bias", "<f8"), ("anchor_a", "<f8", 2), ("anchor_b", "<f8", 2),
    ("prm", "<f8", 4), ("acc", "<f8", 2)], align=True)
SCENE_MAGIC = 0x31454E4543535043

SPACE_PARAMS = np.dtype([
    ("gravity", "<f8", 2), ("damping", "<f8"), ("idle_speed_threshold", "<f8"), ("sleep_time_threshold", "<f8"),
    ("collision_slop", "<f8"), ("collision_bias", "<f8"), ("collision_persistence", "<u4"), ("iterations", "<i4")], align=True)
BODY_DESC = np.dtype([
    ("p", "<f8", 2), ("v", "<f8", 2), ("f", "<f8", 2), ("a", "<f8"), ("w", "<f8"), ("t", "<f8"), ("rot", "<f8", 2),
    ("m", "<f8"), ("i", "<f8"), ("cog", "<f8", 2), ("v_bias", "<f8", 2), ("w_bias", "<f8"), ("idle_time", "<f8"),
    ("type", "<i4"), ("space", "<i4"), ("sleeping", "<i4"), ("sleep_group", "<i4"), ("custom", "<i4"), ("pad_", "<i4")], align=True)
SHAPE_DESC = np.dtype([
    ("type", "<i4"), ("body", "<i4"), ("hashid", "<u4"), ("sensor", "<i4"), ("categories", "<u4"), ("mask", "<u4"),
    ("group", "<u8"), ("collision_type", "<u8"), ("e", "<f8"), ("u", "<f8"), ("surface_v", "<f8", 2), ("r", "<f8"),
    ("a", "<f8", 2), ("b", "<f8", 2), ("a_tangent", "<f8", 2), ("b_tangent", "<f8", 2),
    ("n_verts", "<i4"), ("vert_offset", "<i4")], align=True)
JOINT_DESC = np.dtype([
    ("type", "<i4"), ("a", "<i4"), ("b", "<i4"), ("collide_bodies", "<i4"), ("max_force", "<f8"), ("error_bias", "<f8"),
    ("max_bias", "<f8"), ("anchor_a", "<f8", 2), ("anchor_b", "<f8", 2), ("prm", "<f8", 4), ("acc", "<f8", 2)], align=True)
BODY_STATE = np.dtype([
    ("p", "<f8", 2), ("v", "<f8", 2), ("a", "<f8"), ("w", "<f8"), ("rot", "<f8", 2), ("idle_time", "<f8"),
    ("sleeping", "<i4"), ("sleep_group", "<i4")], align=True)
CONTACT = np.dtype([
    ("r1", "<f8", 2), ("r2", "<f8", 2), ("n_mass", "<f8"), ("t_mass", "<f8"), ("bounce", "<f8"), ("bias", "<f8"),
    ("jn_acc", "<f8"), ("jt_acc", "<f8"), ("j_bias", "<f8"), ("hash", "<u8")], align=True)
ARBITER = np.dtype([
    ("shape_a", "<i4"), ("shape_b", "<i4"), ("body_a", "<i4"), ("body_b", "<i4"), ("count", "<i4"), ("state", "<i4"),
    ("stamp", "<u4"), ("active", "<i4"), ("record", "<i4"), ("pad", "<i4"), ("n", "<f8", 2), ("e", "<f8"), ("u", "<f8"), ("surface_vr", "<f8", 2),
    ("contacts", CONTACT, 2)], align=True)
JOINT_STATE = np.dtype([("acc", "<f8", 2), ("impulse", "<f8"), ("aux", "<f8")], align=True)
STATS = np.dtype([
    ("steps", "<u8"), ("n_bodies", "<u4"), ("n_awake", "<u4"), ("n_shapes", "<u4"), ("n_joints", "<u4"), ("n_pairs", "<u4"),
    ("n_arbiters", "<u4"), ("n_contacts", "<u4"), ("n_cached", "<u4"), ("n_colours", "<u4"), ("overflow", "<u4"),
    ("kinetic_energy", "<f8"), ("max_penetration", "<f8"), ("n_row_solves", "<u4"), ("n_row_idle", "<u4")], align=True)


class EngineError(RuntimeError):
    pass


_lib_cache = {}


def load_engine(path=None):
    """Load libcpb200.so (fails loudly when it has not been built)."""
    path = path or ENGINE_LIB
    if path in _lib_cache:
        return _lib_cache[path]
    if not os.path.exists(path):
        raise EngineError("CUDA engine %s is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                          "There is no CPU fallback." % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    lib.cpb200_world_create.restype = vp
    lib.cpb200_world_create.argtypes = [ci, ci]
    lib.cpb200_world_destroy.argtypes = [vp]
    lib.cpb200_last_error.restype = C.c_char_p
    lib.cpb200_device_available.restype = ci
    lib.cpb200_world_set_space_params.argtypes = [vp, ci, vp]
    lib.cpb200_world_set_bodies.argtypes = [vp, ci, vp]
    lib.cpb200_world_set_shapes.argtypes = [vp, ci, vp, ci, vp]
    lib.cpb200_world_set_joints.argtypes = [vp, ci, vp]
    lib.cpb200_world_update_bodies.argtypes = [vp, ci, ci, vp]
    lib.cpb200_world_set_body_forces.argtypes = [vp, ci, ci, vp]
    lib.cpb200_host_alloc.restype = vp
    lib.cpb200_host_alloc.argtypes = [C.c_size_t]
    lib.cpb200_host_free.restype = None
    lib.cpb200_host_free.argtypes = [vp]
    lib.cpb200_world_reserve.argtypes = [vp, ci, ci]
    lib.cpb200_world_step.argtypes = [vp, cd]
    lib.cpb200_world_sync.argtypes = [vp]
    lib.cpb200_world_time_steps.argtypes = [vp, cd, ci, vp]
    lib.cpb200_launch_count.restype = C.c_ulonglong
    lib.cpb200_world_get_bodies.argtypes = [vp, ci, ci, vp]
    lib.cpb200_world_get_shape_bbs.argtypes = [vp, ci, ci, vp]
    lib.cpb200_world_get_arbiters.argtypes = [vp, ci, vp, ci]
    lib.cpb200_world_get_joints.argtypes = [vp, ci, ci, vp]
    lib.cpb200_world_get_stats.argtypes = [vp, vp]
    lib.cpb200_world_get_pairs.restype = C.c_long
    lib.cpb200_world_get_pairs.argtypes = [vp, C.c_long, vp]
    lib.cpb200_world_set_solver_mode.argtypes = [vp, ci]
    lib.cpb200_world_set_arbiter_order.argtypes = [vp, ci, vp]
    lib.cpb200_world_set_joint_order.argtypes = [vp, ci, vp]
    lib.cpb200_world_set_solver_grid.argtypes = [vp, ci]
    lib.cpb200_world_collide_pair.argtypes = [vp, ci, ci, vp]
    lib.cpb200_world_get_stage_times.argtypes = [vp, ci, vp]
    lib.cpb200_stage_name.restype = C.c_char_p
    lib.cpb200_stage_name.argtypes = [ci]
    lib.cpb200_world_set_profiling.argtypes = [vp, ci]
    lib.cpb200_world_get_solver_profile.argtypes = [vp, vp]
    lib.cpb200_world_step_collide.argtypes = [vp, cd]
    lib.cpb200_world_set_graph.argtypes = [vp, ci]
    lib.cpb200_world_bind_io.argtypes = [vp, vp, vp]
    lib.cpb200_world_append_bodies.argtypes = [vp, ci, vp]
    lib.cpb200_world_append_shapes.argtypes = [vp, ci, vp, ci, vp]
    lib.cpb200_world_append_joints.argtypes = [vp, ci, vp]
    lib.cpb200_world_remove_shape.argtypes = [vp, ci]
    lib.cpb200_world_remove_body.argtypes = [vp, ci]
    lib.cpb200_world_remove_joint.argtypes = [vp, ci]
    lib.cpb200_world_get_graph_stats.argtypes = [vp, vp]
    lib.cpb200_world_graph_error.restype = C.c_char_p
    lib.cpb200_world_graph_error.argtypes = [vp]
    lib.cpb200_world_step_presolve.argtypes = [vp]
    lib.cpb200_world_step_finish.argtypes = [vp]
    lib.cpb200_world_get_body_solver_state.argtypes = [vp, ci, ci, vp]
    lib.cpb200_world_get_joint_solver_state.argtypes = [vp, ci, ci, vp]
    lib.cpb200_world_set_solver_variant.argtypes = [vp, ci]
    lib.cpb200_world_get_solver_order.restype = C.c_long
    lib.cpb200_world_get_solver_order.argtypes = [vp, C.c_long, vp]
    lib.cpb200_world_get_solver_path.argtypes = [vp]
    lib.cpb200_world_get_colour_starts.argtypes = [vp, vp]
    _lib_cache[path] = lib
    return lib


# ---- scene blobs -------------------------------------------------------------------

class Scene:
    """Parsed view of a scene blob (cpb_scene.h)."""

    def __init__(self, blob):
        buf = np.frombuffer(bytes(blob), dtype=np.uint8)
        self.blob = bytes(blob)
        self.header = np.frombuffer(buf[:SCENE_HEADER.itemsize].tobytes(), dtype=SCENE_HEADER)[0]
        if int(self.header["magic"]) != SCENE_MAGIC:
            raise ValueError("not a scene blob")
        off = SCENE_HEADER.itemsize
        nb, ns, nv, nj = (int(self.header[k]) for k in ("n_bodies", "n_shapes", "n_verts", "n_joints"))
        self.bodies = np.frombuffer(buf[off:off + nb * SCENE_BODY.itemsize].tobytes(), dtype=SCENE_BODY); off += nb * SCENE_BODY.itemsize
        self.shapes = np.frombuffer(buf[off:off + ns * SCENE_SHAPE.itemsize].tobytes(), dtype=SCENE_SHAPE); off += ns * SCENE_SHAPE.itemsize
        self.verts = np.frombuffer(buf[off:off + nv * 16].tobytes(), dtype="<f8").reshape(nv, 2); off += nv * 16
        self.joints = np.frombuffer(buf[off:off + nj * SCENE_JOINT.itemsize].tobytes(), dtype=SCENE_JOINT); off += nj * SCENE_JOINT.itemsize
        assert off == len(buf), (off, len(buf))

    @property
    def dt(self):
        return float(self.header["timestep"])

    @staticmethod
    def build(header, bodies, shapes, verts, joints):
        header = np.array(header, dtype=SCENE_HEADER).reshape(())
        header = header.copy()
        header["magic"] = SCENE_MAGIC
        header["n_bodies"], header["n_shapes"] = len(bodies), len(shapes)
        header["n_verts"], header["n_joints"] = len(verts), len(joints)
        parts = [header.tobytes(), np.ascontiguousarray(bodies, dtype=SCENE_BODY).tobytes(),
                 np.ascontiguousarray(shapes, dtype=SCENE_SHAPE).tobytes(),
                 np.ascontiguousarray(verts, dtype="<f8").tobytes(),
                 np.ascontiguousarray(joints, dtype=SCENE_JOINT).tobytes()]
        return Scene(b"".join(parts))

    def n_dynamic(self):
        return int(np.count_nonzero(self.bodies["type"] != 2))


def scene_descs(scene, space=0, body_base=0, vert_base=0, hashid_base=0):
    """Scene records -> C-ABI descriptors (what the C99 host layer produces from its structs)."""
    sb, ss, sj = scene.bodies, scene.shapes, scene.joints
    bd = np.zeros(len(sb), dtype=BODY_DESC)
    for k in ("p", "v", "f", "a", "w", "t", "m", "i", "cog", "type"):
        bd[k] = sb[k]
    bd["cog"][sb["type"] != 0] = 0.0      # the scene format gives only dynamic bodies a centre of gravity (scene_io.c:70-74)
    # the scene's p is what cpBodySetPosition receives, i.e. the body's ORIGIN; the loader sets the centre of gravity
    # first and the angle last, so body->p (what the engine integrates) = p + cog, unrotated (cpBody.c:241-249)
    bd["p"] = sb["p"] + bd["cog"]
    # libm cos/sin (what cpBodySetAngle -> SetTransform uses, cpBody.c:347-357), not numpy's SIMD variants
    bd["rot"][:, 0] = 1.0
    for i in np.nonzero(sb["a"])[0]:
        bd["rot"][i, 0] = math.cos(sb["a"][i])
        bd["rot"][i, 1] = math.sin(sb["a"][i])
    bd["space"] = space
    bd["sleep_group"] = -1
    static = sb["type"] == 2
    bd["idle_time"][static] = np.inf
    sd = np.zeros(len(ss), dtype=SHAPE_DESC)
    for k in ("type", "sensor", "categories", "mask", "group", "collision_type", "e", "u", "surface_v", "r", "a", "b",
              "a_tangent", "b_tangent", "n_verts"):
        sd[k] = ss[k]
    sd["body"] = ss["body"] + body_base
    sd["vert_offset"] = ss["vert_offset"] + vert_base
    sd["hashid"] = np.arange(len(ss), dtype=np.uint32) + hashid_base
    jd = np.zeros(len(sj), dtype=JOINT_DESC)
    for k in ("type", "collide_bodies", "max_force", "error_bias", "max_bias", "anchor_a", "anchor_b", "prm", "acc"):
        jd[k] = sj[k]
    jd["a"] = sj["a"] + body_base
    jd["b"] = sj["b"] + body_base
    return bd, sd, jd


def scene_params(scene):
    h = scene.header
    p = np.zeros((), dtype=SPACE_PARAMS)
    for k in ("gravity", "damping", "idle_speed_threshold", "sleep_time_threshold", "collision_slop", "collision_bias",
              "collision_persistence", "iterations"):
        p[k] = h[k]
    return p


class World:
    """A device world: one or many independent spaces stepped together."""

    def __init__(self, n_spaces=1, device=0, lib_path=None):
        self.lib = load_engine(lib_path)
        self.w = self.lib.cpb200_world_create(device, n_spaces)
        if not self.w:
            raise EngineError(self.lib.cpb200_last_error().decode())
        self.n_spaces = n_spaces
        self.n_bodies = self.n_shapes = self.n_joints = 0

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(self.lib.cpb200_last_error().decode())
        return rc

    def close(self):
        for ptr in getattr(self, "_pinned", []):
            self.lib.cpb200_host_free(ptr)
        self._pinned = []
        if self.w:
            self.lib.cpb200_world_destroy(self.w)
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- upload
    def set_space_params(self, space, params):
        params = np.array(params, dtype=SPACE_PARAMS).reshape(())
        self._ck(self.lib.cpb200_world_set_space_params(self.w, space, params.ctypes.data))

    def set_bodies(self, bd):
        bd = np.ascontiguousarray(bd, dtype=BODY_DESC)
        self._ck(self.lib.cpb200_world_set_bodies(self.w, len(bd), bd.ctypes.data))
        self.n_bodies = len(bd)

    def update_bodies(self, first, bd):
        bd = np.ascontiguousarray(bd, dtype=BODY_DESC)
        self._ck(self.lib.cpb200_world_update_bodies(self.w, first, len(bd), bd.ctypes.data))

    def set_body_forces(self, first, fxyt):
        """Per-step host input: (f.x, f.y, torque) rows for bodies [first, first + len(fxyt))."""
        f = np.ascontiguousarray(fxyt, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.cpb200_world_set_body_forces(self.w, int(first), len(f), f.ctypes.data))

    def pinned_array(self, n, dtype):
        """A numpy array over page-locked host memory (cpb200_host_alloc): transfers from/to it are one DMA at link
        speed.  The memory lives as long as the world."""
        dtype = np.dtype(dtype)
        nbytes = max(1, int(n) * dtype.itemsize)
        ptr = self.lib.cpb200_host_alloc(nbytes)
        if not ptr:
            raise EngineError("cpb200_host_alloc(%d) failed" % nbytes)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_char * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def bind_io(self, forces=None, state_out=None):
        """Bind page-locked per-step I/O buffers (pinned_array): forces[n][3] in, state_out[2][n][3] out, every step."""
        fp = forces.ctypes.data if forces is not None else None
        sp = state_out.ctypes.data if state_out is not None else None
        self._ck(self.lib.cpb200_world_bind_io(self.w, fp, sp))
        self._io = (forces, state_out)

    def bodies_into(self, out):
        """Read-back into a caller-owned BODY_STATE array (no allocation in the step loop)."""
        self._ck(self.lib.cpb200_world_get_bodies(self.w, 0, self.n_bodies, out.ctypes.data))
        return out

    def set_shapes(self, sd, verts):
        sd = np.ascontiguousarray(sd, dtype=SHAPE_DESC)
        verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 2)
        self._ck(self.lib.cpb200_world_set_shapes(self.w, len(sd), sd.ctypes.data, len(verts), verts.ctypes.data))
        self.n_shapes = len(sd)

    def set_joints(self, jd):
        jd = np.ascontiguousarray(jd, dtype=JOINT_DESC)
        self._ck(self.lib.cpb200_world_set_joints(self.w, len(jd), jd.ctypes.data))
        self.n_joints = len(jd)

    def append_bodies(self, bd):
        """Returns False when the arrays' slack is used up (nothing changed; re-upload with set_bodies)."""
        bd = np.ascontiguousarray(bd, dtype=BODY_DESC)
        rc = self._ck(self.lib.cpb200_world_append_bodies(self.w, len(bd), bd.ctypes.data))
        if rc == 0:
            self.n_bodies += len(bd)
        return rc == 0

    def append_shapes(self, sd, verts=None):
        sd = np.ascontiguousarray(sd, dtype=SHAPE_DESC)
        verts = np.ascontiguousarray(verts if verts is not None else np.zeros((0, 2)), dtype=np.float64).reshape(-1, 2)
        rc = self._ck(self.lib.cpb200_world_append_shapes(self.w, len(sd), sd.ctypes.data, len(verts), verts.ctypes.data))
        if rc == 0:
            self.n_shapes += len(sd)
        return rc == 0

    def append_joints(self, jd):
        jd = np.ascontiguousarray(jd, dtype=JOINT_DESC)
        rc = self._ck(self.lib.cpb200_world_append_joints(self.w, len(jd), jd.ctypes.data))
        if rc == 0:
            self.n_joints += len(jd)
        return rc == 0

    def remove_shape(self, index):
        rc = self._ck(self.lib.cpb200_world_remove_shape(self.w, int(index)))
        if rc == 0:
            self.n_shapes -= 1
        return rc == 0

    def remove_body(self, index):
        rc = self._ck(self.lib.cpb200_world_remove_body(self.w, int(index)))
        if rc == 0:
            self.n_bodies -= 1
        return rc == 0

    def remove_joint(self, index):
        rc = self._ck(self.lib.cpb200_world_remove_joint(self.w, int(index)))
        if rc == 0:
            self.n_joints -= 1
        return rc == 0

    def reserve(self, max_pairs=0, max_arbiters=0):
        self._ck(self.lib.cpb200_world_reserve(self.w, int(max_pairs), int(max_arbiters)))

    def load_scene(self, scene):
        """Single-space convenience: upload one scene."""
        self.load_scenes([scene])

    def load_scenes(self, scenes):
        """Upload len(scenes) == n_spaces scenes, one per space, concatenated."""
        assert len(scenes) == self.n_spaces
        bds, sds, jds, vts = [], [], [], []
        bb = vb = hb = 0
        for k, sc in enumerate(scenes):
            self.set_space_params(k, scene_params(sc))
            bd, sd, jd = scene_descs(sc, space=k, body_base=bb, vert_base=vb, hashid_base=hb)
            bds.append(bd); sds.append(sd); jds.append(jd); vts.append(sc.verts)
            bb += len(bd); vb += len(sc.verts); hb += len(sd)
        self.set_bodies(np.concatenate(bds))
        self.set_shapes(np.concatenate(sds), np.concatenate(vts) if vb else np.zeros((0, 2)))
        self.set_joints(np.concatenate(jds))

    # -- step
    def step(self, dt, n=1):
        for _ in range(n):
            self._ck(self.lib.cpb200_world_step(self.w, float(dt)))

    def time_steps(self, dt, n):
        """n steps; returns device milliseconds (CUDA events on the engine's stream)."""
        ms = C.c_float(0.0)
        self._ck(self.lib.cpb200_world_time_steps(self.w, float(dt), int(n), C.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        return int(self.lib.cpb200_launch_count())

    def sync(self):
        self._ck(self.lib.cpb200_world_sync(self.w))

    # -- read-back
    def bodies(self):
        out = np.zeros(self.n_bodies, dtype=BODY_STATE)
        self._ck(self.lib.cpb200_world_get_bodies(self.w, 0, self.n_bodies, out.ctypes.data))
        return out

    def shape_bbs(self):
        out = np.zeros((self.n_shapes, 4))
        self._ck(self.lib.cpb200_world_get_shape_bbs(self.w, 0, self.n_shapes, out.ctypes.data))
        return out

    def arbiters(self, active_only=True):
        n = self._ck(self.lib.cpb200_world_get_arbiters(self.w, 0, None, int(active_only)))
        out = np.zeros(max(n, 1), dtype=ARBITER)
        n = self._ck(self.lib.cpb200_world_get_arbiters(self.w, len(out), out.ctypes.data, int(active_only)))
        return out[:n]

    def joints(self):
        out = np.zeros(max(self.n_joints, 1), dtype=JOINT_STATE)
        self._ck(self.lib.cpb200_world_get_joints(self.w, 0, self.n_joints, out.ctypes.data))
        return out[:self.n_joints]

    def stats(self):
        out = np.zeros((), dtype=STATS)
        self._ck(self.lib.cpb200_world_get_stats(self.w, out.ctypes.data))
        return {k: out[k].item() for k in STATS.names}

    def pairs(self):
        n = self._ck(self.lib.cpb200_world_get_pairs(self.w, 0, None))
        out = np.zeros(max(n, 1), dtype=np.uint64)
        n = self._ck(self.lib.cpb200_world_get_pairs(self.w, len(out), out.ctypes.data))
        return out[:n]

    def set_solver_mode(self, mode):
        self._ck(self.lib.cpb200_world_set_solver_mode(self.w, int(mode)))

    def set_arbiter_order(self, order):
        order = np.ascontiguousarray(order, dtype=np.uint64)
        self._ck(self.lib.cpb200_world_set_arbiter_order(self.w, len(order), order.ctypes.data))

    def set_solver_grid(self, blocks):
        self._ck(self.lib.cpb200_world_set_solver_grid(self.w, int(blocks)))

    def set_joint_order(self, order):
        order = np.ascontiguousarray(order, dtype=np.int32)
        self._ck(self.lib.cpb200_world_set_joint_order(self.w, len(order), order.ctypes.data))

    def collide_pair(self, a, b):
        out = np.zeros(13)
        n = self._ck(self.lib.cpb200_world_collide_pair(self.w, int(a), int(b), out.ctypes.data))
        return n, out

    # -- validation hooks for the production solver order (include/cpb200.h)
    JOINT_SOLVER_ROW = 28

    def step_collide(self, dt):
        self._ck(self.lib.cpb200_world_step_collide(self.w, float(dt)))

    def step_presolve(self):
        self._ck(self.lib.cpb200_world_step_presolve(self.w))

    def step_finish(self):
        self._ck(self.lib.cpb200_world_step_finish(self.w))

    def body_solver_state(self):
        """[n][8] = v.x v.y w m_inv v_bias.x v_bias.y w_bias i_inv, exactly as the solver kernels hold them."""
        out = np.zeros((self.n_bodies, 8))
        self._ck(self.lib.cpb200_world_get_body_solver_state(self.w, 0, self.n_bodies, out.ctypes.data))
        return out

    def joint_solver_state(self):
        out = np.zeros((max(self.n_joints, 1), self.JOINT_SOLVER_ROW))
        self._ck(self.lib.cpb200_world_get_joint_solver_state(self.w, 0, self.n_joints, out.ctypes.data))
        return out[:self.n_joints]

    def set_solver_variant(self, variant):
        self._ck(self.lib.cpb200_world_set_solver_variant(self.w, int(variant)))

    def solver_order(self):
        n = self._ck(self.lib.cpb200_world_get_solver_order(self.w, 0, None))
        out = np.zeros(max(n, 1), dtype=np.int64)
        n = self._ck(self.lib.cpb200_world_get_solver_order(self.w, len(out), out.ctypes.data))
        return out[:n]

    def colour_sizes(self):
        """(rows per colour, joints per colour) of the last world-wide coloured step."""
        out = np.zeros((2, 65), dtype=np.int32)
        self._ck(self.lib.cpb200_world_get_colour_starts(self.w, out.ctypes.data))
        return np.diff(out[0]), np.diff(out[1])

    def solver_path(self):
        return int(self.lib.cpb200_world_get_solver_path(self.w))

    def set_graph(self, on):
        self._ck(self.lib.cpb200_world_set_graph(self.w, int(bool(on))))

    def graph_stats(self):
        out = np.zeros(2, dtype=np.uint64)
        self._ck(self.lib.cpb200_world_get_graph_stats(self.w, out.ctypes.data))
        return {"captures": int(out[0]), "replays": int(out[1]), "error": self.lib.cpb200_world_graph_error(self.w).decode()}

    def set_profiling(self, on):
        self._ck(self.lib.cpb200_world_set_profiling(self.w, int(bool(on))))

    def solver_profile(self):
        buf = np.zeros(6)
        self._ck(self.lib.cpb200_world_get_solver_profile(self.w, buf.ctypes.data))
        return dict(zip(['colour_us', 'rows_us', 'warm_us', 'iterate_us', 'rounds', 'recoloured'], buf.tolist()))

    def stage_times(self):
        buf = np.zeros(32, dtype=np.float32)
        n = self.lib.cpb200_world_get_stage_times(self.w, len(buf), buf.ctypes.data)
        return {self.lib.cpb200_stage_name(i).decode(): float(buf[i]) for i in range(n)}
