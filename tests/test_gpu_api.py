"""The drop-in boundary end to end: host objects built through the public Chipmunk2D C API
(include/chipmunk/chipmunk.h, implemented by chipmunk2d_b200/host/*.c) -> cpSpaceStep on the device
-> public getters.  scenes/scene_io.c is the same translation unit the oracle links against the
reference, so these tests read like the reference's own usage."""
import os
import subprocess

import numpy as np
import pytest

from chipmunk2d_b200.api import load_scene_lib
from chipmunk2d_b200.build import LIB
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import golden_scene
from oracle.ref import SceneSpace

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,steps", [("SimpleTerrainCircles_1000", 120), ("ComplexTerrainHexagons_1000", 120),
                                         ("PyramidStack", 700), ("Chains", 300)])
def test_api_path_equals_engine_path(name, steps):
    """cpSpaceNew/AddBody/AddShape/AddConstraint/cpSpaceStep/getters give bit-identical state to feeding
    the same scene straight into the C ABI: the host layer adds nothing and loses nothing."""
    sc = golden_scene(name)
    api = SceneSpace(load_scene_lib(), sc.blob)
    eng = World(1)
    eng.load_scene(sc)
    for _ in range(steps):
        api.step(sc.dt)
        eng.step(sc.dt)
    eng.sync()
    a = api.bodies()
    e = eng.bodies()
    assert np.array_equal(a[1:, 0:2], e["p"][1:])
    assert np.array_equal(a[1:, 2:4], e["v"][1:])
    assert np.array_equal(a[1:, 4], e["a"][1:]) and np.array_equal(a[1:, 5], e["w"][1:])
    assert np.array_equal(a[1:, 8].astype(int), e["sleeping"][1:])
    assert np.array_equal(api.shape_bbs(), eng.shape_bbs())
    # contact graph through cpBodyEachArbiter + cpArbiterGetContactPointSet
    arbs = api.arbiters()
    dev = eng.arbiters()
    assert len(arbs) == len(dev)
    api.free()


@pytest.mark.parametrize("seed", range(4))
def test_api_path_equals_engine_path_on_random_scenes(seed):
    """The same check on scenes with kinematic bodies, sensors, filters, multi-shape bodies and all ten joint
    classes (tests/test_gpu_fuzz.py), sleeping enabled."""
    from tests.test_gpu_fuzz import random_scene
    sc = random_scene(4000 + seed, n_bodies=60, sleep=0.5)
    api = SceneSpace(load_scene_lib(), sc.blob)
    eng = World(1)
    eng.load_scene(sc)
    for _ in range(250):
        api.step(sc.dt)
        eng.step(sc.dt)
    eng.sync()
    a = api.bodies()
    e = eng.bodies()
    # cpBodyGetPosition is the body's origin: p - rot(cog) (cpBody.c:368-372), the engine reports p
    cog = np.where((sc.bodies["type"] == 0)[:, None], sc.bodies["cog"], 0.0)
    origin = np.stack([e["p"][:, 0] - (cog[:, 0]*e["rot"][:, 0] - cog[:, 1]*e["rot"][:, 1]),
                       e["p"][:, 1] - (cog[:, 0]*e["rot"][:, 1] + cog[:, 1]*e["rot"][:, 0])], axis=1)
    assert np.array_equal(a[1:, 0:2], origin[1:]) and np.array_equal(a[1:, 2:4], e["v"][1:])
    assert np.array_equal(a[1:, 4], e["a"][1:]) and np.array_equal(a[1:, 5], e["w"][1:])
    assert np.array_equal(a[1:, 8].astype(int), e["sleeping"][1:])
    assert np.array_equal(api.shape_bbs(), eng.shape_bbs())
    api.free()


def test_api_hasty_space_is_the_same_device_path():
    sc = golden_scene("SimpleTerrainCircles_100")
    a = SceneSpace(load_scene_lib(), sc.blob)
    h = SceneSpace(load_scene_lib(), sc.blob, hasty=True, threads=2)
    a.step(sc.dt, 100)
    h.step(sc.dt, 100)
    assert np.array_equal(a.bodies(), h.bodies(), equal_nan=True)


def test_api_shapes_collide_matches_reference(ref):
    sc = golden_scene("SimpleTerrainBoxes_100")
    api = SceneSpace(load_scene_lib(), sc.blob)
    rs = ref.load(sc.blob)
    hits = 0
    for a in range(47, 80):
        for b in range(a + 1, 90):
            n_ref, out_ref = rs.shapes_collide(a, b)
            n_dev, out_dev = api.shapes_collide(a, b)
            assert n_ref == n_dev
            if n_ref:
                hits += 1
                assert np.allclose(out_ref, out_dev, rtol=1e-9, atol=1e-9)
    assert hits >= 0


def test_known_answers_of_the_reference_suite(tmp_path):
    """tests/c/known_answers.c (restated XCTest cases) linked against the drop-in."""
    exe = str(tmp_path / "known_answers_b200")
    subprocess.check_call(["gcc", "-O1", "-w", "-o", exe, os.path.join(ROOT, "tests/c/known_answers.c"),
                           "-I", os.path.join(ROOT, "include"), "-L", LIB, "-lchipmunk_b200", "-Wl,-rpath," + LIB, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("PASS") == 8


def test_structural_edit_keeps_warm_start():
    """Adding a body mid-run re-uploads everything; cached arbiters are re-pointed (hashid keyed) so the
    pile does not lose its accumulated impulses: a settled stack stays settled."""
    import ctypes as C
    sc = golden_scene("SimpleTerrainCircles_100")
    lib = load_scene_lib()
    api = SceneSpace(lib, sc.blob)
    api.step(sc.dt, 300)
    before = api.bodies()
    cp = C.CDLL(os.path.join(LIB, "libchipmunk_b200.so"), mode=C.RTLD_LOCAL)
    cp.cpBodyNew.restype = C.c_void_p
    cp.cpBodyNew.argtypes = [C.c_double, C.c_double]
    cp.cpSpaceAddBody.restype = C.c_void_p
    cp.cpSpaceAddBody.argtypes = [C.c_void_p, C.c_void_p]

    class V(C.Structure):
        _fields_ = [("x", C.c_double), ("y", C.c_double)]
    cp.cpBodySetPosition.argtypes = [C.c_void_p, V]
    body = cp.cpSpaceAddBody(api.space, cp.cpBodyNew(1.0, 1.0))
    cp.cpBodySetPosition(body, V(1000.0, 1000.0))
    api.step(sc.dt, 2)
    after = api.bodies()
    # velocities of the resting pile stay small: warm-start impulses survived the re-upload
    assert np.nanmax(np.abs(after[1:, 2:4])) < np.nanmax(np.abs(before[1:, 2:4])) + 15.0
    assert np.nanmax(np.abs(after[1:, 0:2] - before[1:, 0:2])) < 1.0


@pytest.mark.parametrize("which", range(11))
def test_collision_handler_semantics_match_reference(ref, which):
    """begin/preSolve return values, cpArbiterIgnore, cpArbiterSet{Restitution,Friction,SurfaceVelocity}, sensors,
    wildcard and default handlers take effect inside the step they are called in (split device step), with the
    reference's callback counts and the reference's trajectory (scenes/scene_io.c cpb_scene_handler_scenario:
    one ball over static geometry, so the solver order cannot matter).
    0 plain, 1 preSolve false, 2 begin false, 3 restitution, 4 conveyor, 5 sensor, 6 one-way platform,
    7 ignore from the 10th preSolve, 8 wildcard handler, 9 default handler, 10 arbiter user data set in begin is
    seen by every later preSolve/postSolve of the pair."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    out = []
    for lib in (ref.scene, load_scene_lib()):
        lib.cpb_scene_handler_scenario.restype = C.c_int
        lib.cpb_scene_handler_scenario.argtypes = [C.c_int, C.c_int, dp]
        o = np.zeros(12)
        assert lib.cpb_scene_handler_scenario(which, 150, o.ctypes.data_as(dp)) == 0
        out.append(o)
    assert np.array_equal(out[0][:4], out[1][:4]), "callback counts (begin, preSolve, postSolve, separate)"
    assert out[0][10] == out[1][10], "step of the first contact"
    assert np.allclose(out[0], out[1], rtol=1e-9, atol=1e-9)
