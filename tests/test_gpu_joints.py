"""All ten joint classes on the device vs the unmodified reference, lock step in the reference's constraint
order (serial validation mode): preStep / applyCachedImpulse / applyImpulse of pin, slide, pivot, groove,
damped spring, damped rotary spring, rotary limit, ratchet, gear and simple motor
(reference src/cp*Joint.c, cpDamped*Spring.c, cpSimpleMotor.c), plus the coloured order as a sanity check."""
import numpy as np
import pytest

from chipmunk2d_b200.engine import World, Scene, SCENE_HEADER, SCENE_BODY, SCENE_SHAPE, SCENE_JOINT
from chipmunk2d_b200.scenes import ERROR_BIAS_DEFAULT, COLLISION_BIAS_DEFAULT
from tests.util import lockstep, rel_err

pytestmark = pytest.mark.gpu


def all_joints_scene(damping=0.9):
    """A chain of 12 free bodies, each consecutive pair tied by a different joint class; two joints go to the
    static body; one body carries a circle resting on a static segment so contacts and joints share bodies."""
    h = np.zeros((), dtype=SCENE_HEADER)
    h["iterations"] = 10; h["collision_persistence"] = 3; h["gravity"] = (0.0, -100.0); h["damping"] = damping
    h["sleep_time_threshold"] = np.inf; h["collision_slop"] = 0.1; h["collision_bias"] = COLLISION_BIAS_DEFAULT; h["timestep"] = 1.0 / 60.0
    nb = 13
    b = np.zeros(nb, dtype=SCENE_BODY)
    b["type"][0] = 2; b["is_space_static"][0] = 1; b["m"][0] = np.inf; b["i"][0] = np.inf
    for i in range(1, nb):
        b["m"][i] = 1.0 + 0.25 * i; b["i"][i] = 15.0 + 2.0 * i
        b["p"][i] = (10.0 * i, 20.0 + 4.0 * (i % 3)); b["v"][i] = (2.0 - 0.3 * i, 1.0 * i); b["w"][i] = 0.2 * i - 1.0; b["a"][i] = 0.05 * i
    types = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 2]
    j = np.zeros(len(types), dtype=SCENE_JOINT)
    j["max_force"] = np.inf; j["max_bias"] = np.inf; j["error_bias"] = ERROR_BIAS_DEFAULT; j["collide_bodies"] = 1
    j["type"] = types
    j["a"] = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0, 0]
    j["b"] = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 1, 12]
    j["anchor_a"] = [(1, 0), (0, 1), (2, 2), (-6, 1), (-1, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (10, 60), (120, 40)]
    j["anchor_b"] = [(-1, 0), (0, -1), (-8.0, 2.5), (1, 0.5), (1, 1), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0, 1), (0, 2)]
    j["prm"][0, 0] = 9.0                                           # pin dist
    j["prm"][1, :2] = (2.0, 9.0)                                   # slide min max
    j["prm"][3, :2] = (6.0, 2.0)                                   # groove: grv_b (anchor_a is grv_a)
    j["prm"][4, :3] = (8.0, 40.0, 0.7)                             # damped spring
    j["prm"][5, :3] = (0.3, 60.0, 1.5)                             # damped rotary spring
    j["prm"][6, :2] = (-0.2, 0.4)                                  # rotary limit
    j["prm"][7, :3] = (0.05, 0.0, 0.3)                             # ratchet: angle, phase, ratchet
    j["prm"][8, :2] = (0.1, 2.0)                                   # gear phase ratio
    j["prm"][9, 0] = 1.5                                           # motor rate
    j["prm"][10, 0] = 45.0
    j["max_force"][1] = 4000.0; j["max_force"][9] = 500.0; j["max_bias"][2] = 50.0
    s = np.zeros(2, dtype=SCENE_SHAPE)
    s["categories"] = 0xFFFFFFFF; s["mask"] = 0xFFFFFFFF; s["e"] = 0.2; s["u"] = 0.7
    s["type"] = [1, 0]; s["body"] = [0, 6]
    s["a"][0] = (-50, 10); s["b"][0] = (300, 10); s["r"][1] = 6.0
    return Scene.build(h, b, s, np.zeros((0, 2)), j)


def test_all_joint_classes_lockstep(ref):
    sc = all_joints_scene()
    rs = ref.load(sc.blob)
    w = World(1)
    w.load_scene(sc)
    w.set_solver_mode(1)
    worst = {"p": 0.0, "v": 0.0, "j": 0.0}

    def check(step, asleep, arbs, hi):
        rb = rs.priv_bodies(); wb = w.bodies()
        worst["p"] = max(worst["p"], rel_err(wb["p"][1:], rb[1:, 0:2]), rel_err(wb["a"][1:], rb[1:, 4]))
        worst["v"] = max(worst["v"], rel_err(wb["v"][1:], rb[1:, 2:4]), rel_err(wb["w"][1:], rb[1:, 5]))
        rj = rs.priv_joints(); wj = w.joints()
        worst["j"] = max(worst["j"], rel_err(wj["impulse"], rj[:, 11]))

    lockstep(rs, w, sc.dt, 60, check, resync_scene=sc)
    assert worst["p"] < 1e-9 and worst["v"] < 1e-9 and worst["j"] < 1e-7, worst


def test_all_joint_classes_free_running(ref):
    sc = all_joints_scene()
    rs = ref.load(sc.blob)
    w = World(1)
    w.load_scene(sc)
    w.set_solver_mode(1)
    lockstep(rs, w, sc.dt, 120)
    rb = rs.priv_bodies(); wb = w.bodies()
    assert rel_err(wb["p"][1:], rb[1:, 0:2]) < 1e-6 and rel_err(wb["v"][1:], rb[1:, 2:4]) < 1e-5


def test_joints_coloured_order_is_stable():
    sc = all_joints_scene(damping=0.5)
    w = World(1)
    w.load_scene(sc)
    w.step(sc.dt, 600)
    w.sync()
    wb = w.bodies()
    assert np.all(np.isfinite(wb["p"])) and np.all(np.abs(wb["p"]) < 1e4)
    js = w.joints()
    assert np.all(np.isfinite(js["impulse"]))
    # the pin to the static body holds: body 1 stays 45 from its anchor at (10, 60)
    d = np.hypot(wb["p"][1, 0] + np.cos(wb["a"][1]) * 0 - np.sin(wb["a"][1]) * 1 - 10.0, wb["p"][1, 1] + np.sin(wb["a"][1]) * 0 + np.cos(wb["a"][1]) * 1 - 60.0)
    assert abs(d - 45.0) < 1.0, d
