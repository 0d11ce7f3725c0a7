"""Randomised parity: scenes nobody tuned -- random mixes of circles (with offsets), fat segments and convex polygons
(3-8 vertices, with and without bevel) on dynamic, kinematic and static bodies, random filters (groups, category
masks), sensors, restitution / friction / surface velocities, every joint class between random body pairs with
collideBodies on and off -- stepped side by side with the unmodified reference from identical body state every step
(device in the reference's solver order): pair sets bit-exact, p / a / v / w within 1e-9."""
import numpy as np
import pytest

from chipmunk2d_b200.engine import World, Scene, SCENE_HEADER, SCENE_BODY, SCENE_SHAPE, SCENE_JOINT
from chipmunk2d_b200.scenes import ERROR_BIAS_DEFAULT, COLLISION_BIAS_DEFAULT, hull_order, moment_for_poly
from tests.util import lockstep, rel_err

pytestmark = pytest.mark.gpu


def random_scene(seed, n_bodies=40, sleep=np.inf):
    rng = np.random.default_rng(seed)
    h = np.zeros((), dtype=SCENE_HEADER)
    h["iterations"] = int(rng.integers(4, 12)); h["collision_persistence"] = 3
    h["gravity"] = (float(rng.uniform(-20, 20)), float(rng.uniform(-150, -50))); h["damping"] = float(rng.uniform(0.8, 1.0))
    h["sleep_time_threshold"] = sleep; h["idle_speed_threshold"] = 0.0
    h["collision_slop"] = float(rng.choice([0.1, 0.5])); h["collision_bias"] = COLLISION_BIAS_DEFAULT
    h["timestep"] = float(rng.choice([1.0 / 60.0, 1.0 / 120.0]))
    nb = n_bodies + 1
    b = np.zeros(nb, dtype=SCENE_BODY)
    b["type"][0] = 2; b["is_space_static"][0] = 1; b["m"][0] = np.inf; b["i"][0] = np.inf
    shapes, verts = [], []

    def add_shape(body, kind, e, u):
        s = np.zeros((), dtype=SCENE_SHAPE)
        s["body"] = body; s["categories"] = 0xFFFFFFFF; s["mask"] = 0xFFFFFFFF
        s["e"] = e; s["u"] = u
        if rng.random() < 0.15:
            s["surface_v"] = rng.uniform(-30, 30, size=2)
        if rng.random() < 0.2:
            s["group"] = int(rng.integers(1, 4))
        if rng.random() < 0.2:
            s["categories"] = int(rng.integers(1, 8)); s["mask"] = int(rng.integers(1, 8))
        if rng.random() < 0.08:
            s["sensor"] = 1
        if kind == 0:
            s["type"] = 0; s["r"] = float(rng.uniform(2.0, 7.0)); s["a"] = rng.uniform(-2, 2, size=2) * (rng.random() < 0.5)
        elif kind == 1:
            s["type"] = 1; s["r"] = float(rng.choice([0.0, 1.0, 2.5]))
            s["a"] = rng.uniform(-8, 8, size=2); s["b"] = s["a"] + rng.uniform(4, 12) * np.array([np.cos(t := rng.uniform(0, 6.28)), np.sin(t)])
        else:
            k = int(rng.integers(3, 9))
            ang = np.sort(rng.uniform(0, 2 * np.pi, size=k))
            rad = rng.uniform(3.0, 7.0, size=k)
            pts = hull_order(np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=1))
            s["type"] = 2; s["r"] = float(rng.choice([0.0, 0.5, 1.0])); s["n_verts"] = len(pts); s["vert_offset"] = sum(len(v) for v in verts)
            verts.append(pts)
        shapes.append(s)

    # static terrain: a bowl of segments
    for x0, y0, x1, y1 in ((-120, 0, 120, 0), (-120, 0, -150, 80), (120, 0, 150, 80), (-60, 0, -20, 12), (20, 12, 60, 0)):
        s = np.zeros((), dtype=SCENE_SHAPE)
        s["type"] = 1; s["body"] = 0; s["categories"] = 0xFFFFFFFF; s["mask"] = 0xFFFFFFFF; s["e"] = 0.3; s["u"] = 0.8
        s["a"] = (x0, y0); s["b"] = (x1, y1); s["r"] = float(rng.choice([0.0, 1.0]))
        shapes.append(s)
    for i in range(1, nb):
        kin = rng.random() < 0.08
        b["type"][i] = 1 if kin else 0
        b["m"][i] = np.inf if kin else float(rng.uniform(0.5, 5.0)); b["i"][i] = np.inf if kin else float(rng.uniform(10.0, 120.0))
        b["p"][i] = (float(rng.uniform(-100, 100)), float(rng.uniform(8, 90)))
        b["v"][i] = rng.uniform(-15, 15, size=2) * (0.2 if kin else 1.0); b["w"][i] = float(rng.uniform(-2, 2)); b["a"][i] = float(rng.uniform(-3, 3))
        if rng.random() < 0.2 and not kin:    # the scene format gives only dynamic bodies a centre of gravity (scene_io.c)
            b["cog"][i] = rng.uniform(-1, 1, size=2)
        for _ in range(1 if rng.random() < 0.8 else 2):
            add_shape(i, int(rng.integers(0, 3)), float(rng.choice([0.0, 0.3, 0.9])), float(rng.uniform(0.0, 1.0)))
    nj = n_bodies // 3
    j = np.zeros(nj, dtype=SCENE_JOINT)
    j["max_force"] = np.inf; j["max_bias"] = np.inf; j["error_bias"] = ERROR_BIAS_DEFAULT
    for q in range(nj):
        a, c = rng.choice(np.arange(0, nb), size=2, replace=False)
        if b["type"][a] != 0 and b["type"][c] != 0:
            a = int(np.flatnonzero(b["type"] == 0)[0])       # at least one dynamic end
        t = int(rng.integers(0, 10))
        j["type"][q] = t; j["a"][q] = a; j["b"][q] = c; j["collide_bodies"][q] = int(rng.random() < 0.5)
        j["anchor_a"][q] = rng.uniform(-3, 3, size=2); j["anchor_b"][q] = rng.uniform(-3, 3, size=2)
        d = float(np.linalg.norm(b["p"][a] - b["p"][c]))
        if t == 0: j["prm"][q, 0] = d
        elif t == 1: j["prm"][q, :2] = (0.5 * d, 1.2 * d)
        elif t == 3: j["prm"][q, :2] = j["anchor_a"][q] + rng.uniform(4, 10, size=2)
        elif t == 4: j["prm"][q, :3] = (d, float(rng.uniform(20, 200)), float(rng.uniform(0.2, 5.0)))
        elif t == 5: j["prm"][q, :3] = (float(rng.uniform(-1, 1)), float(rng.uniform(50, 500)), float(rng.uniform(0.5, 5.0)))
        elif t == 6: j["prm"][q, :2] = (-0.5, 0.7)
        elif t == 7: j["prm"][q, :3] = (0.0, 0.0, 0.4)
        elif t == 8: j["prm"][q, :2] = (0.2, float(rng.choice([1.0, 2.0, -1.5])))
        elif t == 9: j["prm"][q, 0] = float(rng.uniform(-3, 3))
        if rng.random() < 0.3: j["max_force"][q] = float(rng.uniform(500, 5000))
        if rng.random() < 0.2: j["max_bias"][q] = float(rng.uniform(10, 100))
    v = np.concatenate(verts) if verts else np.zeros((0, 2))
    return Scene.build(h, b, np.array(shapes, dtype=SCENE_SHAPE), v, j)


@pytest.mark.parametrize("seed", range(12))
def test_random_scene_one_step_parity(ref, seed):
    sc = random_scene(1000 + seed)
    rs = ref.load(sc.blob)
    w = World(1)
    w.load_scene(sc)
    w.set_solver_mode(1)
    worst = {"p": 0.0, "v": 0.0, "pairs_bad": 0, "pairs": 0, "arbs": 0}

    def check(step, asleep, arbs, hi):
        pr, pw = rs.pairs(asleep), w.pairs()
        worst["pairs"] += len(pr); worst["arbs"] += len(arbs)
        if not np.array_equal(pr, pw):
            worst["pairs_bad"] += 1
        rb = rs.priv_bodies(); wb = w.bodies()
        worst["p"] = max(worst["p"], rel_err(wb["p"][1:], rb[1:, 0:2]), rel_err(wb["a"][1:], rb[1:, 4]))
        worst["v"] = max(worst["v"], rel_err(wb["v"][1:], rb[1:, 2:4]), rel_err(wb["w"][1:], rb[1:, 5]))

    lockstep(rs, w, sc.dt, 50, check, resync_scene=sc)
    assert worst["pairs_bad"] == 0, worst
    assert worst["pairs"] > 0 and worst["arbs"] > 0
    assert worst["p"] < 1e-9 and worst["v"] < 1e-9, worst
    rs.space = None


@pytest.mark.parametrize("seed", range(4))
def test_random_scene_coloured_solver_is_deterministic_and_sane(seed):
    sc = random_scene(2000 + seed, n_bodies=80, sleep=0.5)
    out = []
    for grid in (0, 5):
        w = World(1); w.load_scene(sc)
        if grid:
            w.set_solver_grid(grid)
        w.step(sc.dt, 300); w.sync()
        out.append(w.bodies())
        assert w.stats()["overflow"] == 0
    assert np.array_equal(out[0]["p"], out[1]["p"]) and np.array_equal(out[0]["v"], out[1]["v"])
    assert np.all(np.isfinite(out[0]["p"])) and np.all(np.isfinite(out[0]["v"]))
