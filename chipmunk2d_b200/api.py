"""ctypes access to the drop-in Chipmunk2D API build (lib/libchipmunk_b200.so) through the scene loader
(lib/libscene_b200.so = scenes/scene_io.c, the same translation unit the oracle side links against the
reference).  Used by tests and bench for the end-to-end path: host objects -> cpSpaceStep -> getters."""
import ctypes as C
import os

import numpy as np

from .engine import LIB_DIR, EngineError

API_LIB = os.path.join(LIB_DIR, "libchipmunk_b200.so")
SCENE_LIB = os.path.join(LIB_DIR, "libscene_b200.so")

_cache = {}


PUB_BODY_ROW = 10  # cpb_scene_get_bodies
PUB_ARB_ROW = 16   # cpb_scene_get_arbiters

_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


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


def load_scene_lib(path=None):
    """Returns the ctypes handle of libscene_b200.so with the scene_io.c entry points declared."""
    path = path or SCENE_LIB
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise EngineError("%s is missing: run __graft_entry__.build()" % path)
    lib = bind_scene_api(C.CDLL(path, mode=C.RTLD_LOCAL))
    _cache[path] = lib
    return lib


def load_api_lib(path=None):
    path = path or API_LIB
    if not os.path.exists(path):
        raise EngineError("%s is missing: run __graft_entry__.build()" % path)
    return C.CDLL(path, mode=C.RTLD_LOCAL)
