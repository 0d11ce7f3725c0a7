"""ctypes front-end for the CPU checkers under oracle/ -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (chipmunk2d_b200) never does.

`Ref` wraps oracle/_ref/libchipmunk_ref.so (the unmodified reference compiled from
/root/reference, plus oracle/ref_probe.c) and oracle/_ref/libscene_ref.so
(chipmunk2d_b200/scenes/scene_io.c linked against the reference).
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

BODY_ROW = 24     # refp_get_bodies
SHAPE_ROW = 10    # refp_get_shapes
ARB_ROW = 36      # refp_get_arbiters
JOINT_ROW = 12    # refp_get_joints
PUB_BODY_ROW = 10  # cpb_scene_get_bodies
PUB_ARB_ROW = 16   # cpb_scene_get_arbiters

_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def available():
    return os.path.exists(os.path.join(REF_DIR, "libchipmunk_ref.so")) and \
        os.path.exists(os.path.join(REF_DIR, "libscene_ref.so"))


def bind_scene_api(lib):
    """Declare the scene_io.c entry points on a ctypes library (used for both builds)."""
    lib.cpb_scene_load.restype = C.c_void_p
    lib.cpb_scene_load.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.cpb_scene_free.restype = None
    lib.cpb_scene_free.argtypes = [C.c_void_p, C.c_int]
    lib.cpb_scene_step.restype = None
    lib.cpb_scene_step.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
    lib.cpb_scene_time_steps.restype = C.c_double
    lib.cpb_scene_time_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
    lib.cpb_scene_get_bodies.restype = None
    lib.cpb_scene_get_bodies.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.cpb_scene_get_shape_bbs.restype = None
    lib.cpb_scene_get_shape_bbs.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.cpb_scene_get_arbiters.restype = C.c_int
    lib.cpb_scene_get_arbiters.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.cpb_scene_shapes_collide.restype = C.c_int
    lib.cpb_scene_shapes_collide.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.cpb_scene_e2e_steps.restype = C.c_double
    lib.cpb_scene_e2e_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int]
    d, vp, ci, u64, u32 = C.c_double, C.c_void_p, C.c_int, C.c_uint64, C.c_uint32
    lib.cpb_scene_point_query.restype = ci
    lib.cpb_scene_point_query.argtypes = [vp, d, d, d, u64, u32, u32, ci, _dp]
    lib.cpb_scene_point_query_nearest.restype = ci
    lib.cpb_scene_point_query_nearest.argtypes = [vp, d, d, d, u64, u32, u32, _dp]
    lib.cpb_scene_segment_query.restype = ci
    lib.cpb_scene_segment_query.argtypes = [vp, d, d, d, d, d, u64, u32, u32, ci, _dp]
    lib.cpb_scene_segment_query_first.restype = ci
    lib.cpb_scene_segment_query_first.argtypes = [vp, d, d, d, d, d, u64, u32, u32, _dp]
    lib.cpb_scene_bb_query.restype = ci
    lib.cpb_scene_bb_query.argtypes = [vp, d, d, d, d, u64, u32, u32, ci, _dp]
    lib.cpb_scene_shape_point_query.restype = ci
    lib.cpb_scene_shape_point_query.argtypes = [vp, ci, d, d, _dp]
    lib.cpb_scene_shape_segment_query.restype = ci
    lib.cpb_scene_shape_segment_query.argtypes = [vp, ci, d, d, d, d, d, _dp]
    lib.cpb_scene_shape_query.restype = ci
    lib.cpb_scene_shape_query.argtypes = [vp, ci, d, d, d, d, d, d, ci, _dp, C.POINTER(ci)]
    return lib


class SceneSpace:
    """A cpSpace instantiated from a scene blob in one of the two libraries,
    driven only through scene_io.c (public API)."""

    def __init__(self, scene_lib, blob, hasty=False, threads=0):
        self.lib = scene_lib
        self.blob = np.frombuffer(bytes(blob), dtype=np.uint8).copy()
        hdr = np.frombuffer(self.blob[:32].tobytes(), dtype=np.int32)
        self.n_bodies, self.n_shapes, self.n_verts, self.n_joints = (int(x) for x in hdr[2:6])
        self.hasty = int(bool(hasty))
        self.space = self.lib.cpb_scene_load(self.blob.ctypes.data, self.hasty, int(threads))
        if not self.space:
            raise RuntimeError("cpb_scene_load failed")

    def step(self, dt, n=1):
        self.lib.cpb_scene_step(self.space, dt, n, self.hasty)

    def time_steps(self, dt, n):
        return self.lib.cpb_scene_time_steps(self.space, dt, n, self.hasty)

    def e2e_steps(self, dt, n, force=(0.0, 0.0)):
        """n steps with per-step host writes (forces) and host reads (positions); returns (seconds, positions)."""
        out = np.zeros((self.n_bodies, 2))
        t = self.lib.cpb_scene_e2e_steps(self.space, dt, n, self.n_bodies, _p(out), force[0], force[1], self.hasty)
        return t, out

    def bodies(self):
        out = np.full((self.n_bodies, PUB_BODY_ROW), np.nan)
        self.lib.cpb_scene_get_bodies(self.space, self.n_bodies, _p(out))
        return out

    def shape_bbs(self):
        out = np.full((self.n_shapes, 4), np.nan)
        self.lib.cpb_scene_get_shape_bbs(self.space, self.n_shapes, _p(out))
        return out

    def arbiters(self, cap=None):
        cap = cap or max(16, 8 * self.n_shapes)
        out = np.zeros((cap, PUB_ARB_ROW))
        n = self.lib.cpb_scene_get_arbiters(self.space, cap, _p(out))
        if n > cap:
            return self.arbiters(cap=n)
        return out[:n]

    def shapes_collide(self, ia, ib):
        out = np.zeros(13)
        n = self.lib.cpb_scene_shapes_collide(self.space, ia, ib, _p(out))
        return n, out

    # -- space queries (public cpSpace*Query API); filter = (group, categories, mask)
    ALL = (0, 0xffffffff, 0xffffffff)

    def _rows(self, call, stride, cap=256):
        while True:
            out = np.zeros((cap, stride))
            n = call(cap, _p(out))
            if n <= cap:
                return out[:n]
            cap = n

    def point_query(self, p, max_dist, filt=ALL):
        return self._rows(lambda cap, o: self.lib.cpb_scene_point_query(self.space, p[0], p[1], max_dist, filt[0], filt[1], filt[2], cap, o), 6)

    def point_query_nearest(self, p, max_dist, filt=ALL):
        out = np.zeros(6)
        hit = self.lib.cpb_scene_point_query_nearest(self.space, p[0], p[1], max_dist, filt[0], filt[1], filt[2], _p(out))
        return bool(hit), out

    def segment_query(self, a, b, radius=0.0, filt=ALL):
        return self._rows(lambda cap, o: self.lib.cpb_scene_segment_query(self.space, a[0], a[1], b[0], b[1], radius, filt[0], filt[1], filt[2], cap, o), 6)

    def segment_query_first(self, a, b, radius=0.0, filt=ALL):
        out = np.zeros(6)
        hit = self.lib.cpb_scene_segment_query_first(self.space, a[0], a[1], b[0], b[1], radius, filt[0], filt[1], filt[2], _p(out))
        return bool(hit), out

    def bb_query(self, bb, filt=ALL):
        return self._rows(lambda cap, o: self.lib.cpb_scene_bb_query(self.space, bb[0], bb[1], bb[2], bb[3], filt[0], filt[1], filt[2], cap, o), 1)[:, 0].astype(int)

    def shape_point_query(self, tag, p):
        out = np.zeros(6)
        rc = self.lib.cpb_scene_shape_point_query(self.space, tag, p[0], p[1], _p(out))
        return rc, out

    def shape_segment_query(self, tag, a, b, radius=0.0):
        out = np.zeros(6)
        rc = self.lib.cpb_scene_shape_segment_query(self.space, tag, a[0], a[1], b[0], b[1], radius, _p(out))
        return rc, out

    def shape_query(self, kind, pos, angle=0.0, w=0.0, h=0.0, radius=0.0):
        any_ = C.c_int(0)
        rows = self._rows(lambda cap, o: self.lib.cpb_scene_shape_query(self.space, kind, pos[0], pos[1], angle, w, h, radius, cap, o, C.byref(any_)), 14)
        return rows, bool(any_.value)

    def free(self):
        if self.space:
            self.lib.cpb_scene_free(self.space, self.hasty)
            self.space = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Ref:
    """The unmodified reference + probe."""

    def __init__(self):
        if not available():
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
        self.cp = C.CDLL(os.path.join(REF_DIR, "libchipmunk_ref.so"), mode=C.RTLD_LOCAL)
        self.scene = bind_scene_api(C.CDLL(os.path.join(REF_DIR, "libscene_ref.so"), mode=C.RTLD_LOCAL))
        cp = self.cp
        cp.refp_demo_count.restype = C.c_int
        cp.refp_demo_name.restype = C.c_char_p
        cp.refp_demo_name.argtypes = [C.c_int]
        cp.refp_demo_build.restype = C.c_void_p
        cp.refp_demo_build.argtypes = [C.c_char_p]
        cp.refp_demo_timestep.restype = C.c_double
        cp.refp_demo_timestep.argtypes = [C.c_char_p]
        cp.refp_scene_dump.restype = C.c_size_t
        cp.refp_scene_dump.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]
        cp.refp_get_bodies.argtypes = [C.c_void_p, C.c_int, _dp]
        cp.refp_get_shapes.argtypes = [C.c_void_p, C.c_int, _dp]
        cp.refp_get_poly_planes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), _dp]
        cp.refp_get_arbiters.restype = C.c_int
        cp.refp_get_arbiters.argtypes = [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_uint)]
        cp.refp_get_joints.argtypes = [C.c_void_p, C.c_int, _dp]
        cp.refp_pairs_bruteforce.restype = C.c_long
        cp.refp_pairs_bruteforce.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_uint64)]
        cp.refp_step.argtypes = [C.c_void_p, C.c_double, C.c_int]
        cp.refp_get_constraint_order.restype = C.c_int
        cp.refp_get_constraint_order.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        cp.refp_space_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        cp.cpSpaceStep.argtypes = [C.c_void_p, C.c_double]
        cp.refp_install_order_hook.argtypes = [C.c_void_p]
        cp.refp_set_solver_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        cp.refp_order_hook_stats.argtypes = [C.POINTER(C.c_int)]

    def demo_names(self):
        return [self.cp.refp_demo_name(i).decode() for i in range(self.cp.refp_demo_count())]

    def demo_scene(self, name):
        """Build a demo scene with the reference's own demo code and flatten it.
        Returns (blob bytes, timestep).  The demo space itself is leaked (tiny)."""
        space = self.cp.refp_demo_build(name.encode())
        if not space:
            raise KeyError(name)
        dt = self.cp.refp_demo_timestep(name.encode())
        need = self.cp.refp_scene_dump(space, dt, None, 0)
        buf = np.zeros(need, dtype=np.uint8)
        self.cp.refp_scene_dump(space, dt, buf.ctypes.data, need)
        return buf.tobytes(), dt

    def demo_space(self, name):
        space = self.cp.refp_demo_build(name.encode())
        if not space:
            raise KeyError(name)
        dt = self.cp.refp_demo_timestep(name.encode())
        # tag userData
        self.cp.refp_scene_dump(space, dt, None, 0)
        return space, dt

    def load(self, blob, hasty=False, threads=0):
        return RefSpace(self, blob, hasty, threads)


class RefSpace(SceneSpace):
    """SceneSpace in the reference, plus private-state probes."""

    def __init__(self, ref, blob, hasty=False, threads=0):
        super().__init__(ref.scene, blob, hasty, threads)
        self.ref = ref

    def priv_bodies(self):
        out = np.full((self.n_bodies, BODY_ROW), np.nan)
        self.ref.cp.refp_get_bodies(self.space, self.n_bodies, _p(out))
        return out

    def priv_shapes(self):
        out = np.zeros((self.n_shapes, SHAPE_ROW))
        self.ref.cp.refp_get_shapes(self.space, self.n_shapes, _p(out))
        return out

    def poly_planes(self, vert_offsets):
        vo = np.ascontiguousarray(vert_offsets, dtype=np.int32)
        out = np.zeros((max(self.n_verts, 1), 4))
        self.ref.cp.refp_get_poly_planes(self.space, self.n_shapes, vo.ctypes.data_as(C.POINTER(C.c_int)), _p(out))
        return out[:self.n_verts]

    def priv_arbiters(self):
        cap = max(16, 8 * self.n_shapes)
        out = np.zeros((cap, ARB_ROW))
        hi = np.zeros(2 * cap, dtype=np.uint32)
        n = self.ref.cp.refp_get_arbiters(self.space, cap, _p(out), hi.ctypes.data_as(C.POINTER(C.c_uint)))
        assert n <= cap
        return out[:n], hi[:2 * n].reshape(n, 2)

    def priv_joints(self):
        out = np.zeros((max(self.n_joints, 1), JOINT_ROW))
        self.ref.cp.refp_get_joints(self.space, self.n_joints, _p(out))
        return out[:self.n_joints]

    def constraint_order(self):
        out = np.zeros(max(self.n_joints, 1), dtype=np.int32)
        n = self.ref.cp.refp_get_constraint_order(self.space, len(out), out.ctypes.data_as(C.POINTER(C.c_int)))
        return out[:n]

    def pairs(self, asleep=None):
        cap = 64 * max(self.n_shapes, 16)
        out = np.zeros(cap, dtype=np.uint64)
        ap = None
        if asleep is not None:
            asleep = np.ascontiguousarray(asleep, dtype=np.uint8)
            ap = asleep.ctypes.data
        n = self.ref.cp.refp_pairs_bruteforce(self.space, ap, cap, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert n <= cap
        return out[:n]

    def install_order_hook(self):
        """Route every body's velocity_func through the probe's wrapper so that set_solver_order can permute
        space->arbiters / space->constraints right before the solver loop (ref_probe.c, 'solver-order hook')."""
        self.ref.cp.refp_install_order_hook(self.space)

    def set_solver_order(self, pairs, hash0=None, joints=()):
        """Order for the NEXT step: pairs = (shape a << 32 | shape b) per arbiter, hash0 = hash of the contact to
        visit first per arbiter, joints = scene joint indices."""
        pairs = np.ascontiguousarray(pairs, dtype=np.uint64)
        h = None if hash0 is None else np.ascontiguousarray(hash0, dtype=np.uint64)
        j = np.ascontiguousarray(joints, dtype=np.int32)
        self.ref.cp.refp_set_solver_order(self.space, len(pairs), pairs.ctypes.data, None if h is None else h.ctypes.data,
                                          len(j), j.ctypes.data if len(j) else None, self.n_joints)

    def order_hook_stats(self):
        out = (C.c_int * 2)()
        self.ref.cp.refp_order_hook_stats(out)
        return {"applied": out[0], "unmatched": out[1]}

    def counts(self):
        out = (C.c_int * 8)()
        self.ref.cp.refp_space_counts(self.space, out)
        keys = ["dynamic_bodies", "static_bodies", "arbiters", "constraints", "stamp", "sleeping_components", "contacts"]
        return dict(zip(keys, list(out)))
