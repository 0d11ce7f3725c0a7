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

_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def available():
    return os.path.exists(os.path.join(REF_DIR, "libchipmunk_ref.so")) and \
        os.path.exists(os.path.join(REF_DIR, "libscene_ref.so"))


from chipmunk2d_b200.api import SceneSpace, bind_scene_api, PUB_BODY_ROW, PUB_ARB_ROW  # noqa: E402,F401  (generic scene_io.c wrapper)


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
