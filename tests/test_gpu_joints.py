"""All ten joint classes on the device vs the unmodified reference, lock step in the reference's constraint
order (serial validation mode): preStep / applyCachedImpulse / applyImpulse of pin, slide, pivot, groove,
damped spring, damped rotary spring, rotary limit, ratchet, gear and simple motor
(reference src/cp*Joint.c, cpDamped*Spring.c, cpSimpleMotor.c), plus the coloured order as a sanity check."""
import numpy as np
import pytest

from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import all_joints_scene
from tests.util import lockstep, rel_err

pytestmark = pytest.mark.gpu


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
