"""f4: structural additions by range (cpb200_world_append_bodies / shapes / joints; SURVEY.md 8f rank 4,
cpSpaceAddBody / AddShape / AddConstraint, cpSpace.c:417-474).  An appended world must evolve bit for bit like a world
that was rebuilt from scratch with the same objects (the round-1 path: download everything, re-upload everything),
and appending must cost microseconds, not a re-upload of the world."""
import time

import numpy as np
import pytest

from chipmunk2d_b200.engine import World, scene_descs, BODY_DESC, SHAPE_DESC, JOINT_DESC
from chipmunk2d_b200.scenes import circle_pile, golden_scene, mixed_drop

pytestmark = pytest.mark.gpu


def newcomers(w, k, kind, y0=900.0):
    """k new dynamic bodies dropped above the scene, one shape each (circle / box), every second pair pinned together."""
    nb, ns = w.n_bodies, w.n_shapes
    bd = np.zeros(k, dtype=BODY_DESC)
    bd["m"] = 2.0; bd["i"] = 40.0; bd["rot"][:, 0] = 1.0; bd["sleep_group"] = -1
    bd["p"][:, 0] = 200.0 + 23.0 * np.arange(k); bd["p"][:, 1] = y0; bd["v"][:, 1] = -50.0
    sd = np.zeros(k, dtype=SHAPE_DESC)
    sd["body"] = nb + np.arange(k); sd["hashid"] = 1000000 + ns + np.arange(k)
    sd["categories"] = 0xFFFFFFFF; sd["mask"] = 0xFFFFFFFF; sd["e"] = 0.1; sd["u"] = 0.7
    verts = np.zeros((0, 2))
    if kind == "circle":
        sd["type"] = 0; sd["r"] = 6.0
    else:
        sd["type"] = 2; sd["r"] = 0.5; sd["n_verts"] = 4; sd["vert_offset"] = 4 * np.arange(k)
        verts = np.tile(np.array([[4.0, -4.0], [4.0, 4.0], [-4.0, 4.0], [-4.0, -4.0]]), (k, 1))
    jd = np.zeros(k // 2, dtype=JOINT_DESC)
    jd["type"] = 0; jd["a"] = nb + 2 * np.arange(k // 2); jd["b"] = jd["a"] + 1
    jd["max_force"] = np.inf; jd["max_bias"] = np.inf; jd["error_bias"] = 0.9 ** 60; jd["collide_bodies"] = 0
    jd["prm"][:, 0] = 23.0
    return bd, sd, verts, jd


def reupload_everything(w, sc, extra):
    """The round-1 way on the same world: read everything back, upload everything again with the newcomers behind it
    (cpb200_world_set_bodies / set_shapes / set_joints; the arbiter cache is re-pointed by set_shapes)."""
    from chipmunk2d_b200.engine import scene_descs
    bd0, sd0, jd0 = scene_descs(sc)
    st = w.bodies()
    bias = w.body_solver_state()
    bd0 = bd0.copy()
    for k in ("p", "v", "a", "w", "rot", "idle_time", "sleeping", "sleep_group"):
        bd0[k] = st[k]
    bd0["v_bias"] = bias[:, 4:6]; bd0["w_bias"] = bias[:, 6]
    jd0 = jd0.copy()
    if len(jd0):
        jd0["acc"] = w.joints()["acc"]
    bd, sd, verts, jd = extra
    sd = sd.copy(); sd["vert_offset"] += len(sc.verts)
    w.set_bodies(np.concatenate([bd0, bd]))
    w.set_shapes(np.concatenate([sd0, sd]), np.concatenate([sc.verts, verts]) if (len(verts) + len(sc.verts)) else np.zeros((0, 2)))
    w.set_joints(np.concatenate([jd0, jd]))


@pytest.mark.parametrize("name,kind", [("SimpleTerrainCircles_1000", "circle"), ("ComplexTerrainHexagons_1000", "box"), ("mixed6k", "box")])
def test_append_equals_reuploading_the_whole_world(name, kind, monkeypatch):
    sc = mixed_drop(6000) if name == "mixed6k" else golden_scene(name)
    # colours from scratch every step (they depend on stable ids only): a re-upload resets the kept colours, and with them
    # the Gauss-Seidel order, which an append does not -- the comparison must not see that difference
    monkeypatch.setenv("CPB200_NO_HINTS", "1")
    a, b = World(1), World(1)
    for w in (a, b):
        # (room for every arbiter up front: a re-upload that has to GROW the arbiter buffers drops the cached arbiters --
        # the warm start an append always keeps -- and the two worlds would then legitimately differ)
        w.reserve(max_pairs=400000, max_arbiters=200000)
        w.load_scene(sc)
        w.step(sc.dt, 30)
        w.sync()
    assert np.array_equal(a.bodies()["p"], b.bodies()["p"])
    extra = newcomers(a, 6, kind, y0=(400.0 if name != "mixed6k" else 2500.0))
    bd, sd, verts, jd = extra
    assert a.append_bodies(bd) and a.append_shapes(sd, verts) and a.append_joints(jd)
    reupload_everything(b, sc, extra)
    for s in range(60):
        a.step(sc.dt); b.step(sc.dt)
    a.sync(); b.sync()
    x, y = a.bodies(), b.bodies()
    for k in ("p", "v", "a", "w"):
        assert np.array_equal(x[k], y[k]), (name, k, float(np.max(np.abs(x[k] - y[k]))))
    assert np.array_equal(a.pairs(), b.pairs())
    assert a.stats()["n_joints"] == len(sc.joints) + 3
    assert x["p"][-1, 1] < (400.0 if name != "mixed6k" else 2500.0)       # the newcomers fall


def test_append_keeps_the_warm_start_and_matches_an_upfront_world():
    """Bodies parked far away from the start are equivalent to bodies appended later at the same place: the appended world
    must then match a world that had them from the beginning BIT FOR BIT (same colours, same cached impulses)."""
    sc = golden_scene("ComplexTerrainHexagons_1000")
    w = World(1); w.load_scene(sc)
    w.step(sc.dt, 40); w.sync()
    extra = newcomers(w, 4, "circle", y0=5000.0)
    bd, sd, verts, jd = extra
    bd = bd.copy(); bd["v"][:, 1] = 0.0
    # reference world: the same four bodies present from step 0, static until step 40 (they touch nothing up there)
    bd0, sd0, jd0 = scene_descs(sc)
    parked = bd.copy(); parked["type"] = 2; parked["m"] = np.inf; parked["i"] = np.inf; parked["idle_time"] = np.inf
    from chipmunk2d_b200.engine import scene_params
    ref = World(1); ref.set_space_params(0, scene_params(sc))
    ref.set_bodies(np.concatenate([bd0, parked])); ref.set_shapes(np.concatenate([sd0, sd]), sc.verts); ref.set_joints(jd0)
    ref.step(sc.dt, 40); ref.sync()
    assert np.array_equal(w.bodies()["p"], ref.bodies()["p"][:w.n_bodies])
    assert w.append_bodies(bd) and w.append_shapes(sd, verts)
    g0 = w.graph_stats()
    w.step(sc.dt, 30); ref.step(sc.dt, 30)
    w.sync(); ref.sync()
    # the old world's bodies are unaffected by the append: identical to a world that never re-uploaded anything
    assert np.array_equal(w.bodies()["p"][:len(bd0)], ref.bodies()["p"][:len(bd0)])
    assert np.array_equal(w.bodies()["v"][:len(bd0)], ref.bodies()["v"][:len(bd0)])
    assert w.bodies()["p"][-1, 1] < 5000.0          # and the newcomers fall
    assert w.graph_stats()["replays"] > g0["replays"]


def test_append_costs_microseconds_on_a_large_world():
    """SURVEY 8(f) rank 4 / VERDICT: a 1 M-body space must absorb one added body + shape in < 1 ms of extra time."""
    n = 1000000
    sc = circle_pile(n, dense=True, sleep=0.5)
    w = World(1); w.load_scene(sc)
    w.step(sc.dt, 5); w.sync()

    def timed_steps(k):
        t0 = time.perf_counter(); w.step(sc.dt, k); w.sync(); return (time.perf_counter() - t0) / k
    base = min(timed_steps(4) for _ in range(3))
    costs = []
    for r in range(5):
        bd, sd, verts, jd = newcomers(w, 1, "circle", y0=9000.0 + 20.0 * r)
        t0 = time.perf_counter()
        assert w.append_bodies(bd) and w.append_shapes(sd, verts)
        w.step(sc.dt); w.sync()
        costs.append(time.perf_counter() - t0 - base)
        w.step(sc.dt, 2); w.sync()
    extra_ms = 1000.0 * float(np.median(costs))
    print("append + first step: %.3f ms over a plain step of %.3f ms" % (extra_ms, 1000.0 * base))
    assert extra_ms < 1.0, (costs, base)
    assert w.n_bodies == n + 1 + 5 and w.stats()["overflow"] == 0


def test_append_reports_exhausted_slack():
    sc = golden_scene("SimpleTerrainCircles_100")
    w = World(1); w.load_scene(sc)
    w.step(sc.dt, 3); w.sync()
    bd, sd, verts, jd = newcomers(w, 2000, "circle")
    assert w.append_bodies(bd) is False          # 101 bodies carry 281 slots of slack
    assert w.n_bodies == len(sc.bodies)
    w.step(sc.dt, 3); w.sync()


@pytest.mark.parametrize("name", ["SimpleTerrainCircles_1000", "mixed6k"])
def test_removal_in_place_equals_reuploading_the_world_without_the_object(name, monkeypatch):
    """cpb200_world_remove_shape / remove_joint / remove_body (swap with last + re-pointing on the device) against a
    re-upload of the world without the object, in the same slot order: identical evolution afterwards."""
    from chipmunk2d_b200.engine import scene_descs
    sc = mixed_drop(6000) if name == "mixed6k" else golden_scene(name)
    monkeypatch.setenv("CPB200_NO_HINTS", "1")
    a, b = World(1), World(1)
    for w in (a, b):
        w.reserve(max_pairs=400000, max_arbiters=200000)
        w.load_scene(sc)
        w.step(sc.dt, 40)
        w.sync()
    bd0, sd0, jd0 = scene_descs(sc)
    nb, ns, nj = len(bd0), len(sd0), len(jd0)
    victims = [nb // 2, 7, nb - 1]                                   # a middle body, an early one, the last one
    # host-side model of the registries: current slot -> original index
    body_slot, shape_slot, joint_slot = list(range(nb)), list(range(ns)), list(range(nj))
    for v in victims:
        k = body_slot.index(v)                                       # where that body sits now
        for q in sorted([i for i, j in enumerate(joint_slot) if jd0["a"][j] == v or jd0["b"][j] == v], reverse=True):
            assert a.remove_joint(q); joint_slot[q] = joint_slot[-1]; joint_slot.pop()
        for q in sorted([i for i, s_ in enumerate(shape_slot) if sd0["body"][s_] == v], reverse=True):
            assert a.remove_shape(q); shape_slot[q] = shape_slot[-1]; shape_slot.pop()
        assert a.remove_body(k); body_slot[k] = body_slot[-1]; body_slot.pop()
    assert (a.n_bodies, a.n_shapes, a.n_joints) == (len(body_slot), len(shape_slot), len(joint_slot))
    # world b: the same registries uploaded from scratch (state read back first)
    st = b.bodies(); bias = b.body_solver_state()
    bd = bd0.copy()
    for key in ("p", "v", "a", "w", "rot", "idle_time", "sleeping", "sleep_group"):
        bd[key] = st[key]
    bd["v_bias"] = bias[:, 4:6]; bd["w_bias"] = bias[:, 6]
    new_body = {orig: slot for slot, orig in enumerate(body_slot)}
    bd = bd[body_slot]
    sd = sd0[shape_slot].copy(); sd["body"] = [new_body[int(x)] for x in sd["body"]]
    jd = jd0[joint_slot].copy()
    if nj:
        jd["acc"] = b.joints()["acc"][joint_slot]
        jd["a"] = [new_body[int(x)] for x in jd["a"]]; jd["b"] = [new_body[int(x)] for x in jd["b"]]
    b.set_bodies(bd); b.set_shapes(sd, sc.verts); b.set_joints(jd)
    for s in range(50):
        a.step(sc.dt); b.step(sc.dt)
    a.sync(); b.sync()
    x, y = a.bodies(), b.bodies()
    for key in ("p", "v", "a", "w"):
        assert np.array_equal(x[key], y[key]), (name, key, float(np.max(np.abs(x[key] - y[key]))))
    assert np.array_equal(a.pairs(), b.pairs())
    assert a.stats()["overflow"] == 0
