"""Space queries: the public cpSpace*Query / cpShape*Query API of the drop-in (device scans)
against the unmodified reference, through the same scene_io.c harness compiled against both libraries.

At step 0 both libraries hold bit-identical world caches (the host's libm rotations are uploaded), so the results
must be bit-identical.  After stepping on the device the state is snapshotted into a new scene and loaded into
the reference: the caches then differ by the rounding of sin/cos, and results agree to 1e-9."""
import numpy as np
import pytest

from chipmunk2d_b200.api import load_scene_lib
from chipmunk2d_b200.engine import Scene
from chipmunk2d_b200.scenes import golden_scene
from oracle.ref import SceneSpace, bind_scene_api

pytestmark = pytest.mark.gpu

SCENES = ["SimpleTerrainCircles_100", "SimpleTerrainBoxes_100", "SimpleTerrainHexagons_100", "PyramidStack", "Chains"]


def ours(blob):
    return SceneSpace(bind_scene_api(load_scene_lib()), blob)


def probes(sc, seed):
    """Query geometry spread over the bodies of the scene (deterministic)."""
    rng = np.random.default_rng(seed)
    p = sc.bodies["p"]
    lo, hi = p.min(axis=0) - 30.0, p.max(axis=0) + 30.0
    pts = rng.uniform(lo, hi, size=(24, 2))
    near = p[rng.integers(0, len(p), 8)] + rng.uniform(-6, 6, size=(8, 2))
    return np.concatenate([pts, near]), lo, hi, rng


def thin_ray_crosses(bb, a, b):
    """cpBBIntersectsSegment (cpBB.h:115-149)."""
    tmin, tmax = -np.inf, np.inf
    for lo, hi, s, d in ((bb[0], bb[2], a[0], b[0] - a[0]), (bb[1], bb[3], a[1], b[1] - a[1])):
        if d == 0.0:
            if s < lo or hi < s:
                return False
        else:
            t1, t2 = (lo - s)/d, (hi - s)/d
            tmin, tmax = max(tmin, min(t1, t2)), min(tmax, max(t1, t2))
    return tmin <= tmax and 0.0 <= tmax and tmin <= 1.0


def same_rows(a, b, exact):
    assert a.shape == b.shape, (a.shape, b.shape)
    if exact:
        assert np.array_equal(a, b)
    else:
        assert np.array_equal(a[:, 0], b[:, 0])
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9)


def run_queries(dev, ref, sc, exact, seed=7):
    pts, lo, hi, rng = probes(sc, seed)
    n_hits = 0
    for k, p in enumerate(pts):
        for md in (0.0, 7.5, 40.0):
            a, b = dev.point_query(p, md), ref.point_query(p, md)
            same_rows(a, b, exact)
            n_hits += len(a)
            (ha, ra), (hb, rb) = dev.point_query_nearest(p, md), ref.point_query_nearest(p, md)
            assert ha == hb
            if ha:
                # equal distances (symmetric stacks) may pick different shapes: the distance is what is defined
                assert (ra[3] == rb[3]) if exact else abs(ra[3] - rb[3]) < 1e-9
                if ra[0] == rb[0]:
                    same_rows(ra[None], rb[None], exact)
        q = pts[(k + 5) % len(pts)]
        bbs = dev.shape_bbs()
        for radius in (0.0, 2.0):
            a, b = dev.segment_query(p, q, radius), ref.segment_query(p, q, radius)
            if radius == 0.0:
                same_rows(a, b, exact)
            else:
                # The reference prunes its tree with the THIN ray against node boxes (cpBBTree.c:367-385), so a fat
                # ray loses hits whose boxes the centre line misses -- an artefact of its index.  The device scan
                # reports every geometric hit: a superset, identical on the common shapes.
                common = np.isin(a[:, 0], b[:, 0])
                assert common.sum() == len(b)
                same_rows(a[common], b, exact)
                for row in a[~common]:
                    assert not thin_ray_crosses(bbs[int(row[0])], p, q)
            n_hits += len(a)
            (ha, ra), (hb, rb) = dev.segment_query_first(p, q, radius), ref.segment_query_first(p, q, radius)
            if radius == 0.0:
                assert ha == hb
                if ha:
                    assert (ra[5] == rb[5]) if exact else abs(ra[5] - rb[5]) < 1e-9
                    if ra[0] == rb[0]:
                        same_rows(ra[None], rb[None], exact)
            elif hb:
                assert ha and ra[5] <= rb[5] + 1e-9
        w = rng.uniform(5, 60, size=2)
        bb = (p[0] - w[0], p[1] - w[1], p[0] + w[0], p[1] + w[1])
        assert np.array_equal(dev.bb_query(bb), ref.bb_query(bb))
    assert n_hits > 50
    # filters: a category mask that matches nothing, and the groups used by the scene
    assert len(dev.point_query(pts[0], 1e9, (0, 0, 0))) == 0
    allhits = dev.point_query(pts[0], 1e9)
    assert len(allhits) == len(sc.shapes)
    groups = np.unique(sc.shapes["group"])
    for g in groups[:3]:
        same_rows(dev.point_query(pts[0], 1e9, (int(g), 0xffffffff, 0xffffffff)), ref.point_query(pts[0], 1e9, (int(g), 0xffffffff, 0xffffffff)), exact)
    return n_hits


@pytest.mark.parametrize("name", SCENES)
def test_queries_bit_identical_on_the_loaded_scene(ref, name):
    sc = golden_scene(name)
    dev, rs = ours(sc.blob), ref.load(sc.blob)
    run_queries(dev, rs, sc, exact=True)
    dev.free()
    rs.space = None


@pytest.mark.parametrize("name", ["SimpleTerrainCircles_100", "SimpleTerrainBoxes_100", "PyramidStack"])
def test_queries_after_stepping_match_the_reference_on_the_same_state(ref, name):
    sc = golden_scene(name)
    dev = ours(sc.blob)
    dev.step(sc.dt, 150)
    st = dev.bodies()
    bodies = sc.bodies.copy()
    moved = ~np.isnan(st[:, 0])
    bodies["p"][moved] = st[moved, 0:2]
    bodies["v"][moved] = st[moved, 2:4]
    bodies["a"][moved] = st[moved, 4]
    bodies["w"][moved] = st[moved, 5]
    snap = Scene.build(sc.header, bodies, sc.shapes, sc.verts, sc.joints)
    rs = ref.load(snap.blob)
    run_queries(dev, rs, snap, exact=False, seed=11)
    dev.free()
    rs.space = None


@pytest.mark.parametrize("name", ["SimpleTerrainBoxes_100", "SimpleTerrainHexagons_100", "Chains"])
def test_per_shape_queries(ref, name):
    sc = golden_scene(name)
    dev, rs = ours(sc.blob), ref.load(sc.blob)
    pts, lo, hi, rng = probes(sc, 3)
    tags = rng.integers(0, len(sc.shapes), 40)
    for t, p, q in zip(tags, pts, np.roll(pts, 3, axis=0)):
        (ra, a), (rb, b) = dev.shape_point_query(int(t), p), rs.shape_point_query(int(t), p)
        assert ra == rb == 1 and np.array_equal(a, b)
        for radius in (0.0, 3.0):
            (ha, a), (hb, b) = dev.shape_segment_query(int(t), p, q, radius), rs.shape_segment_query(int(t), p, q, radius)
            assert ha == hb and np.array_equal(a, b)
    dev.free()
    rs.space = None


@pytest.mark.parametrize("name", ["SimpleTerrainCircles_100", "SimpleTerrainBoxes_100", "PyramidStack"])
def test_shape_query_with_a_probe_outside_the_space(ref, name):
    sc = golden_scene(name)
    dev, rs = ours(sc.blob), ref.load(sc.blob)
    pts, lo, hi, rng = probes(sc, 5)
    total = 0
    for k, p in enumerate(pts[-12:]):
        for kind, args in ((0, dict(radius=12.0)), (1, dict(w=30.0, h=-14.0, radius=2.0)), (2, dict(angle=0.3*k, w=25.0, h=18.0, radius=1.0))):
            (a, any_a), (b, any_b) = dev.shape_query(kind, p, **args), rs.shape_query(kind, p, **args)
            assert a.shape == b.shape and np.array_equal(a[:, 0:2], b[:, 0:2])
            # contact points come out of GJK/EPA: 1e-9 like every narrowphase comparison (the probe's rotation is
            # computed by each library's own host code, bit-identical here)
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9)
            assert any_a == any_b
            total += len(a)
    assert total > 5
    dev.free()
    rs.space = None


@pytest.mark.parametrize("seed", range(6))
def test_queries_on_random_scenes(ref, seed):
    """Random polygons (3-8 vertices, bevels), fat segments, offset circles, groups, category masks and sensors
    (tests/test_gpu_fuzz.py): every query kind bit-identical to the reference on the loaded scene."""
    from tests.test_gpu_fuzz import random_scene
    sc = random_scene(3000 + seed, n_bodies=60)
    dev, rs = ours(sc.blob), ref.load(sc.blob)
    run_queries(dev, rs, sc, exact=True, seed=20 + seed)
    pts, lo, hi, rng = probes(sc, 40 + seed)
    for k, p in enumerate(pts[:10]):
        for kind, args in ((0, dict(radius=9.0)), (1, dict(w=25.0, h=11.0, radius=1.5)), (2, dict(angle=0.4*k, w=20.0, h=12.0, radius=0.5))):
            (a, any_a), (b, any_b) = dev.shape_query(kind, p, **args), rs.shape_query(kind, p, **args)
            assert a.shape == b.shape and np.array_equal(a[:, 0:2], b[:, 0:2])
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9)
    dev.free()
    rs.space = None
