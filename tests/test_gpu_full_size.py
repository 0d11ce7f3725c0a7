"""BASELINE configs 4 and 5 at their FULL sizes, through properties that do not need the reference to step them.

The unmodified reference needs 19 s per step at 1 M bodies (and overflows its stack in the sleeping pass beyond 500 k),
so the lockstep tests stop at 50 k bodies (tests/test_gpu_scale_parity.py).  At the benchmark's own sizes the device is
checked against what the algorithm DEFINES, computed independently on the host:

  config 4 (1 M circles, one space, the bench.py workload)
    * broadphase contract (cpSpaceStep.c:219-247): the reported pair set is EXACTLY the set of
      shape pairs whose cached AABBs intersect (cpBBIntersects, closed intervals) minus the QueryReject rules --
      candidates from a k-d tree over the AABB centres, decided on the device's own AABBs, compared as sorted keys;
    * narrowphase (CircleToCircle, cpCollision.c:347-361): the active circle-circle arbiters are EXACTLY the pairs
      with |c2-c1|^2 < (r1+r2)^2, and every contact normal is bit-identical to delta * (1/dist) in IEEE double;
    * solver (cpSpaceStep.c:406-427, cpArbiter.c:441-498): one production step of the world-wide persistent kernel
      replayed through oracle/cp_oracle.c in the device's own order -- 2.9 M arbiters x 10 iterations, bit-identical.
  config 3 (100 k circles / boxes / hexagons with damped springs and pivots)
    * the same pair-set identity on polygon AABBs, and the same bit-for-bit solver replay with its ~10 k joints.
  config 5 (4096 PyramidStack / Chains spaces in one world)
    * spaces are independent (cpSpace.c:119-184 keeps no globals): every one of the 4096 spaces is bit-identical, body
      for body, to the same scene stepped alone in a one-space world (which the lockstep tests pin to the reference).
"""
import os

import numpy as np
import pytest

from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import circle_pile, batched_demo_scenes, mixed_drop
from tests.replay import production_step_replay

pytestmark = pytest.mark.gpu

N_PILE = int(os.environ.get("CPB200_TEST_PILE", 1000000))        # smaller only for the emulator build (tools/emu)
N_BATCH = int(os.environ.get("CPB200_TEST_BATCH", 4096))


@pytest.fixture(scope="module")
def pile():
    sc = circle_pile(N_PILE, dense=True, sleep=0.5)
    w = World(1)
    w.load_scene(sc)
    w.step(sc.dt, 12)                       # the column starts to collapse: contacts open and close at every step
    w.sync()
    yield sc, w
    w.close()


def expected_pairs(sc, bbs, asleep=None):
    """Sorted (min<<32|max) keys of every shape pair with intersecting AABBs and at least one awake dynamic body
    (`asleep` = per-body sleeping flags at collision time, i.e. from before the step)."""
    from scipy.spatial import cKDTree
    shapes = sc.shapes
    static = (shapes["body"] == 0)
    awake = np.ones(len(shapes), dtype=bool) if asleep is None else (asleep[shapes["body"]] == 0)
    awake &= ~static
    dyn = np.nonzero(~static)[0]
    ctr = np.column_stack([(bbs[dyn, 0] + bbs[dyn, 2]) * 0.5, (bbs[dyn, 1] + bbs[dyn, 3]) * 0.5])
    half = max(float(np.max(bbs[dyn, 2] - bbs[dyn, 0])), float(np.max(bbs[dyn, 3] - bbs[dyn, 1]))) * 0.5
    # Chebyshev ball of the largest possible centre distance of two intersecting boxes (+ rounding room): a superset
    cand = cKDTree(ctr).query_pairs(r=2.0 * half * (1.0 + 1e-9) + 1e-9, p=np.inf, output_type="ndarray")
    a, b = dyn[cand[:, 0]], dyn[cand[:, 1]]
    A, Bb = bbs[a], bbs[b]
    hit = (A[:, 0] <= Bb[:, 2]) & (Bb[:, 0] <= A[:, 2]) & (A[:, 1] <= Bb[:, 3]) & (Bb[:, 1] <= A[:, 3])      # cpBBIntersects
    hit &= (shapes["body"][a] != shapes["body"][b]) & (awake[a] | awake[b])
    keys = [(np.minimum(a, b)[hit].astype(np.uint64) << np.uint64(32)) | np.maximum(a, b)[hit].astype(np.uint64)]
    by_x = np.argsort(ctr[:, 0], kind="stable")
    xs = ctr[by_x, 0]
    for s in np.nonzero(static)[0]:          # floor pieces and walls: only the dynamic shapes in their x range can touch
        S = bbs[s]
        d = dyn[by_x[np.searchsorted(xs, S[0] - half - 1e-6):np.searchsorted(xs, S[2] + half + 1e-6, side="right")]]
        D = bbs[d]
        h = (S[0] <= D[:, 2]) & (D[:, 0] <= S[2]) & (S[1] <= D[:, 3]) & (D[:, 1] <= S[3]) & awake[d]
        d = d[h]
        keys.append((np.minimum(d, s).astype(np.uint64) << np.uint64(32)) | np.maximum(d, s).astype(np.uint64))
    return np.sort(np.concatenate(keys))


def test_pile_1m_pair_set_is_exactly_the_intersecting_aabbs(pile):
    sc, w = pile
    st = w.stats()
    assert st["overflow"] == 0 and st["n_awake"] == N_PILE      # nothing can be asleep 12 steps in (threshold 0.5 s)
    got = w.pairs()
    want = expected_pairs(sc, w.shape_bbs())
    assert len(want) > 2 * N_PILE
    assert np.array_equal(got, want), (len(got), len(want))


def test_pile_1m_circle_arbiters_are_exactly_the_touching_pairs(pile):
    sc, w = pile
    shapes = sc.shapes
    arbs = w.arbiters(active_only=True)
    pairs = w.pairs()
    lo = (pairs >> np.uint64(32)).astype(np.int64); hi = (pairs & np.uint64(0xFFFFFFFF)).astype(np.int64)
    cc = (shapes["type"][lo] == 0) & (shapes["type"][hi] == 0)
    lo, hi = lo[cc], hi[cc]
    p = w.bodies()["p"]
    c1, c2 = p[shapes["body"][lo]], p[shapes["body"][hi]]       # circle offset (0,0): the world centre is the body position
    d = c2 - c1
    distsq = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
    mind = shapes["r"][lo] + shapes["r"][hi]
    touching = distsq < mind * mind
    want = np.sort((lo[touching].astype(np.uint64) << np.uint64(32)) | hi[touching].astype(np.uint64))
    a_cc = (shapes["type"][arbs["shape_a"]] == 0) & (shapes["type"][arbs["shape_b"]] == 0)
    A = arbs[a_cc]
    sa, sb = A["shape_a"].astype(np.int64), A["shape_b"].astype(np.int64)
    got_keys = (np.minimum(sa, sb).astype(np.uint64) << np.uint64(32)) | np.maximum(sa, sb).astype(np.uint64)
    order = np.argsort(got_keys)
    assert len(want) > 2 * N_PILE
    assert np.array_equal(got_keys[order], want), (len(got_keys), len(want))
    assert np.all(A["count"] == 1)
    # the contact normal: delta * (1/dist) from shape_a to shape_b, IEEE double, bit for bit
    ca, cb = p[shapes["body"][sa]], p[shapes["body"][sb]]
    dd = cb - ca
    dist = np.sqrt(dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1])
    assert np.all(dist > 0.0)
    n = dd * (1.0 / dist)[:, None]
    assert np.array_equal(A["n"], n)
    # contact points: r1 = (c_a + n*r_a) - p_a, r2 = (c_b + n*(-r_b)) - p_b  (cpCollision.c:358, cpArbiter.c:334-335)
    ra, rb = shapes["r"][sa][:, None], shapes["r"][sb][:, None]
    assert np.array_equal(A["contacts"]["r1"][:, 0], (ca + n * ra) - ca)
    assert np.array_equal(A["contacts"]["r2"][:, 0], (cb + n * (-rb)) - cb)


def test_pile_1m_production_step_replayed_by_the_oracle_bit_for_bit(pile):
    sc, w = pile
    it = int(sc.header["iterations"])
    for _ in range(2):
        r = production_step_replay(w, sc.dt, 1.0, it)
        assert r.path == 1                                       # k_colour_solve, the kernel the bench times
        assert r.n_arbiters > 2 * N_PILE and r.n_items == r.n_arbiters
        assert r.max_rel <= 1e-9 and r.bit_equal, r.detail
    assert w.stats()["overflow"] == 0


def test_pile_1m_asleep_is_frozen_and_wakes_where_it_is_hit():
    """Config 4's "sleeping islands" at full size (cpSpaceComponent.c:234-349): a 20-row pile settles and falls asleep;
    asleep it reports no pairs and does not move by one bit; a body thrown into it wakes the component it touches (the whole pile is ONE contact
    component, and cpBodyActivate wakes components whole, cpSpaceComponent.c:56-91), and from then on the pair set is again exactly the intersecting AABBs with at least one AWAKE body."""
    from chipmunk2d_b200.engine import scene_descs
    n = N_PILE
    sc = circle_pile(n, dense=True, sleep=0.5, columns=max(8, n // 20))
    w = World(1)
    w.load_scene(sc)
    w.step(sc.dt, 700)
    w.sync()
    st = w.stats()
    assert st["overflow"] == 0 and st["n_awake"] == 0, st
    b0 = w.bodies()
    assert np.all(b0["sleeping"][1:] == 1) and len(w.pairs()) == 0
    w.step(sc.dt, 5); w.sync()
    b1 = w.bodies()
    for f in ("p", "v", "a", "w"):
        assert np.array_equal(b0[f], b1[f]), f
    # throw one circle of the top row sideways through its neighbours
    bd, _, _ = scene_descs(sc)
    k = int(np.argmax(b1["p"][1:, 1])) + 1
    d = bd[k:k + 1].copy()
    d["p"] = b1["p"][k]; d["a"] = b1["a"][k]; d["rot"] = b1["rot"][k]; d["v"] = (400.0, -50.0); d["sleeping"] = 0; d["sleep_group"] = -1
    w.update_bodies(k, d)
    w.step(sc.dt, 3); w.sync()
    woke = []
    for s in range(4):
        asleep = w.bodies()["sleeping"]
        w.step(sc.dt); w.sync()
        got = w.pairs()
        want = expected_pairs(sc, w.shape_bbs(), asleep)
        assert len(want) > 0 and np.array_equal(got, want), (s, len(got), len(want))
        woke.append(w.stats()["n_awake"])
    assert woke[-1] > 1, woke                                    # the hit woke more than the thrown body
    assert w.stats()["overflow"] == 0
    w.close()


def test_mixed_100k_pair_set_and_production_step_replay():
    n = int(os.environ.get("CPB200_TEST_MIXED", 100000))
    sc = mixed_drop(n)
    assert len(sc.joints) > n // 20
    w = World(1)
    w.load_scene(sc)
    w.step(sc.dt, 120)                      # bench.py's settle count for this workload: the heap has landed
    w.sync()
    assert np.all(w.bodies()["sleeping"] == 0)
    got = w.pairs()
    want = expected_pairs(sc, w.shape_bbs())
    assert len(want) > n // 4 and np.array_equal(got, want), (len(got), len(want))
    it = int(sc.header["iterations"])
    for _ in range(2):
        r = production_step_replay(w, sc.dt, 1.0, it)
        assert r.path == 1 and r.n_joints == len(sc.joints) and r.n_arbiters > n // 4
        assert r.max_rel <= 1e-9 and r.bit_equal, r.detail
    assert w.stats()["overflow"] == 0
    w.close()


def test_batch_of_4096_spaces_every_space_equals_its_scene_stepped_alone():
    n, steps = N_BATCH, 120
    scenes = batched_demo_scenes(n)
    dt = scenes[0].dt
    alone = []
    for sc in scenes[:2]:
        w1 = World(1); w1.load_scene(sc); w1.step(dt, steps); w1.sync()
        alone.append(w1.bodies()); w1.close()
    w = World(n)
    w.load_scenes(scenes)
    w.step(dt, steps)
    w.sync()
    assert w.solver_path() == 2                                  # the space-local solver, as in the config-5 bench lines
    assert w.stats()["overflow"] == 0
    wb = w.bodies()
    sizes = [len(sc.bodies) for sc in scenes]
    assert sizes[0::2] == [sizes[0]] * (n // 2) and sizes[1::2] == [sizes[1]] * (n // 2)
    stride = sizes[0] + sizes[1]
    for kind in (0, 1):
        first = 0 if kind == 0 else sizes[0]
        for field in ("p", "v", "a", "w"):
            x = wb[field].reshape(n // 2, stride, -1)[:, first:first + sizes[kind]]
            want = alone[kind][field].reshape(1, sizes[kind], -1)
            assert np.array_equal(x, np.broadcast_to(want, x.shape)), (kind, field)
