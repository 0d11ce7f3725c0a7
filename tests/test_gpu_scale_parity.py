"""BASELINE configs 3, 4 and 5 against the unmodified reference at sizes the reference steps in seconds.

  config 4  circle_pile(50 000, dense, sleepTimeThreshold 0.5): the PRODUCTION step (LBVH at 17 Morton bits with ties,
            candidate lists + k_pair_filter, k_collide<0/1>, device arbiter table, islands pass, coloured solver)
            against the reference solving in the device's order: pair sets bit-exact, state 1e-9, every step.
  config 3  mixed_drop(20 000) circles / boxes / hexagons: the same without joints in production order, and WITH its
            damped springs and pivots in the reference's order (serial validation mode).
  config 5  a 64-space PyramidStack / Chains batch against 64 reference spaces (serial order), every space's pair
            set bit-exact and state 1e-9 at every step.
  capacity  exhausted pair / arbiter buffers raise from cpb200_world_sync instead of dropping collisions.
"""
import numpy as np
import pytest

from chipmunk2d_b200.engine import World, EngineError
from chipmunk2d_b200.scenes import circle_pile, mixed_drop, batched_demo_scenes
from tests.util import oracle_body_descs, rel_err, order_keys, lockstep
from tests.test_gpu_production_vs_ref import device_order_for_reference

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.mark.parametrize("name,steps", [("circle_pile_50000", 40), ("mixed_drop_20000_no_joints", 40)])  # 40 + steps
def test_reduced_scale_production_step_equals_reference_in_the_same_order(ref, name, steps):
    sc = circle_pile(50000, dense=True, sleep=0.5) if name == "circle_pile_50000" else mixed_drop(20000, joints=False)
    rs = ref.load(sc.blob)
    rs.install_order_hook()
    w = World(1)
    w.load_scene(sc)
    worst = {"p": 0.0, "v": 0.0, "pairs_bad": 0, "pairs": 0, "unmatched": 0}
    for s in range(steps):
        rb0 = rs.priv_bodies()
        asleep = np.nan_to_num(rb0[:, 19]).astype(np.uint8)
        w.update_bodies(0, oracle_body_descs(rb0, sc, w.bodies()))
        w.step(sc.dt)
        w.sync()
        assert w.solver_path() == 1                      # the world-wide persistent kernel, as at 1 M bodies
        pairs, hash0, joints = device_order_for_reference(w)
        rs.set_solver_order(pairs, hash0, joints)
        rs.step(sc.dt)
        worst["unmatched"] += rs.order_hook_stats()["unmatched"]
        pw = w.pairs()
        worst["pairs"] += len(pw)
        if not np.array_equal(rs.pairs(asleep), pw):
            worst["pairs_bad"] += 1
        rb = rs.priv_bodies(); wb = w.bodies()
        worst["p"] = max(worst["p"], rel_err(wb["p"][1:], rb[1:, 0:2]), rel_err(wb["a"][1:], rb[1:, 4]))
        worst["v"] = max(worst["v"], rel_err(wb["v"][1:], rb[1:, 2:4]), rel_err(wb["w"][1:], rb[1:, 5]))
    st = w.stats()
    assert st["overflow"] == 0
    assert worst["pairs"] > (steps * len(sc.bodies) if name == "circle_pile_50000" else 20000), worst     # the dense pile: ~3 pairs per body at every step
    assert worst["pairs_bad"] == 0 and worst["unmatched"] == 0, worst
    assert worst["p"] < TOL and worst["v"] < TOL, worst
    rs.space = None


def test_mixed_drop_20000_with_springs_and_pivots_serial_lockstep(ref):
    sc = mixed_drop(20000)
    assert len(sc.joints) >= 1900
    rs = ref.load(sc.blob)
    w = World(1)
    w.load_scene(sc)
    w.set_solver_mode(1)
    worst = {"p": 0.0, "v": 0.0, "pairs_bad": 0, "pairs": 0}

    def check(step, asleep, arbs, hi):
        pr, pw = rs.pairs(asleep), w.pairs()
        worst["pairs"] += len(pr)
        if not np.array_equal(pr, pw):
            worst["pairs_bad"] += 1
        rb = rs.priv_bodies(); wb = w.bodies()
        worst["p"] = max(worst["p"], rel_err(wb["p"][1:], rb[1:, 0:2]), rel_err(wb["a"][1:], rb[1:, 4]))
        worst["v"] = max(worst["v"], rel_err(wb["v"][1:], rb[1:, 2:4]), rel_err(wb["w"][1:], rb[1:, 5]))

    lockstep(rs, w, sc.dt, 40, check, resync_scene=sc)
    assert worst["pairs"] > 0 and worst["pairs_bad"] == 0, worst
    assert worst["p"] < TOL and worst["v"] < TOL, worst
    rs.space = None


def test_batch_of_64_spaces_against_64_reference_spaces(ref):
    """Config 5 layout at reduced count: every space of the batch evolves as the reference's own space does."""
    n = 64
    scenes = batched_demo_scenes(n)
    refs = [ref.load(sc.blob) for sc in scenes]
    w = World(n)
    w.load_scenes(scenes)
    w.set_solver_mode(1)
    dt = scenes[0].dt
    body0 = np.cumsum([0] + [len(sc.bodies) for sc in scenes])
    shape0 = np.cumsum([0] + [len(sc.shapes) for sc in scenes])
    joint0 = np.cumsum([0] + [len(sc.joints) for sc in scenes])
    worst = {"p": 0.0, "v": 0.0, "pairs_bad": 0, "pairs": 0, "sleep_bad": 0}
    steps = 300
    for s in range(steps):
        check = (s < 3 or s % 10 == 9 or s > steps - 20)
        dev = w.bodies()
        descs, asleep = [], []
        for k, (rs, sc) in enumerate(zip(refs, scenes)):
            rb = rs.priv_bodies()
            d = oracle_body_descs(rb, sc, dev[body0[k]:body0[k + 1]])
            d["space"] = k
            # sleep groups are world-wide body indices on the device
            descs.append(d)
            asleep.append(np.nan_to_num(rb[:, 19]).astype(np.uint8))
        w.update_bodies(0, np.concatenate(descs))
        order, jorder = [], []
        for k, rs in enumerate(refs):
            rs.step(dt)
            arbs, _hi = rs.priv_arbiters()
            if len(arbs):
                order.append(((arbs[:, 0].astype(np.uint64) + np.uint64(shape0[k])) << np.uint64(32)) | (arbs[:, 1].astype(np.uint64) + np.uint64(shape0[k])))
            if rs.n_joints:
                jorder.append(rs.constraint_order().astype(np.int32) + np.int32(joint0[k]))
        w.set_arbiter_order(np.concatenate(order) if order else np.zeros(0, dtype=np.uint64))
        w.set_joint_order(np.concatenate(jorder) if jorder else np.zeros(0, dtype=np.int32))
        w.step(dt)
        w.sync()
        if not check:
            continue
        pw = w.pairs()
        pr = []
        for k, rs in enumerate(refs):
            p = rs.pairs(asleep[k])
            lo = (p >> np.uint64(32)) + np.uint64(shape0[k]); hi = (p & np.uint64(0xFFFFFFFF)) + np.uint64(shape0[k])
            pr.append((lo << np.uint64(32)) | hi)
        pr = np.sort(np.concatenate(pr))
        worst["pairs"] += len(pr)
        if not np.array_equal(pr, pw):
            worst["pairs_bad"] += 1
        wb = w.bodies()
        for k, rs in enumerate(refs):
            rb = rs.priv_bodies(); b = wb[body0[k]:body0[k + 1]]
            worst["p"] = max(worst["p"], rel_err(b["p"][1:], rb[1:, 0:2]), rel_err(b["a"][1:], rb[1:, 4]))
            worst["v"] = max(worst["v"], rel_err(b["v"][1:], rb[1:, 2:4]), rel_err(b["w"][1:], rb[1:, 5]))
            if not np.array_equal(np.nan_to_num(rb[1:, 19]).astype(int), b["sleeping"][1:]):
                worst["sleep_bad"] += 1
    assert w.stats()["overflow"] == 0
    assert worst["pairs"] > 0 and worst["pairs_bad"] == 0 and worst["sleep_bad"] == 0, worst
    assert worst["p"] < TOL and worst["v"] < TOL, worst


@pytest.mark.parametrize("what", ["pairs", "arbiters"])
def test_exhausted_capacity_raises_instead_of_dropping_collisions(what):
    """'overflow sets a flag that cpb200_world_sync turns into an error -- never silent'.  300 mutually
    overlapping circles make 44 850 pairs; the default capacities for 300 shapes are 5 824 candidates / 3 424 arbiters."""
    from chipmunk2d_b200.engine import Scene, SCENE_SHAPE, SCENE_BODY, SCENE_JOINT
    from chipmunk2d_b200.scenes import _header, _static_body
    n = 300
    b = np.zeros(n + 1, dtype=SCENE_BODY)
    b[0] = _static_body()[0]
    b["m"][1:] = 1.0; b["i"][1:] = 10.0
    b["p"][1:, 0] = np.linspace(0.0, 3.0, n); b["p"][1:, 1] = np.linspace(0.0, 2.0, n)
    s = np.zeros(n, dtype=SCENE_SHAPE)
    s["body"] = np.arange(1, n + 1); s["r"] = 20.0; s["categories"] = 0xFFFFFFFF; s["mask"] = 0xFFFFFFFF
    sc = Scene.build(_header(), b, s, np.zeros((0, 2)), np.zeros(0, dtype=SCENE_JOINT))
    w = World(1)
    if what == "arbiters":
        w.reserve(max_pairs=200000)          # enough pairs, too few arbiter records
    w.load_scene(sc)
    w.step(sc.dt)
    with pytest.raises(EngineError, match="overflow"):
        w.sync()
    # with room for everything the same scene steps and reports every pair
    w2 = World(1)
    w2.reserve(max_pairs=200000, max_arbiters=100000)
    w2.load_scene(sc)
    w2.step(sc.dt)
    w2.sync()
    assert len(w2.pairs()) == n * (n - 1) // 2
