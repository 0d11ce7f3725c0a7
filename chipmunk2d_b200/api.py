"""ctypes access to the drop-in Chipmunk2D API build (lib/libchipmunk_b200.so) through the scene loader
(lib/libscene_b200.so = scenes/scene_io.c, the same translation unit the oracle side links against the
reference).  Used by tests and bench for the end-to-end path: host objects -> cpSpaceStep -> getters."""
import ctypes as C
import os

from .engine import LIB_DIR, EngineError

API_LIB = os.path.join(LIB_DIR, "libchipmunk_b200.so")
SCENE_LIB = os.path.join(LIB_DIR, "libscene_b200.so")

_cache = {}


def load_scene_lib(path=None):
    """Returns the ctypes handle of libscene_b200.so with the scene_io.c entry points declared."""
    path = path or SCENE_LIB
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise EngineError("%s is missing: run __graft_entry__.build()" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    dp = C.POINTER(C.c_double)
    lib.cpb_scene_load.restype = C.c_void_p
    lib.cpb_scene_load.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.cpb_scene_free.restype = None
    lib.cpb_scene_free.argtypes = [C.c_void_p, C.c_int]
    lib.cpb_scene_step.restype = None
    lib.cpb_scene_step.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
    lib.cpb_scene_time_steps.restype = C.c_double
    lib.cpb_scene_time_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
    lib.cpb_scene_get_bodies.restype = None
    lib.cpb_scene_get_bodies.argtypes = [C.c_void_p, C.c_int, dp]
    lib.cpb_scene_get_shape_bbs.restype = None
    lib.cpb_scene_get_shape_bbs.argtypes = [C.c_void_p, C.c_int, dp]
    lib.cpb_scene_get_arbiters.restype = C.c_int
    lib.cpb_scene_get_arbiters.argtypes = [C.c_void_p, C.c_int, dp]
    lib.cpb_scene_shapes_collide.restype = C.c_int
    lib.cpb_scene_shapes_collide.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    lib.cpb_scene_e2e_steps.restype = C.c_double
    lib.cpb_scene_e2e_steps.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, dp, C.c_double, C.c_double, C.c_int]
    _cache[path] = lib
    return lib


def load_api_lib(path=None):
    path = path or API_LIB
    if not os.path.exists(path):
        raise EngineError("%s is missing: run __graft_entry__.build()" % path)
    return C.CDLL(path, mode=C.RTLD_LOCAL)
