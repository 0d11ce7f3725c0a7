"""The oracle itself (no GPU): the unmodified reference compiled under oracle/_ref reproduces the
committed golden vectors bit for bit, passes the restated known-answer tests of its own suite, and the
scene blob round trip (reference demo code -> blob -> public API) is exact."""
import os
import subprocess

import numpy as np
import pytest

from chipmunk2d_b200.scenes import golden_scene, golden_names
from tests.util import golden_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["SimpleTerrainCircles_100", "SimpleTerrainHexagons_100", "PyramidStack", "Chains", "ComplexTerrainHexagons_1000"])
def test_reference_reproduces_golden(ref, name):
    sc = golden_scene(name)
    g = golden_ref(name)
    rs = ref.load(sc.blob)
    done = 0
    for k, s in enumerate(g["steps"]):
        if s > 300:
            break
        rs.step(sc.dt, int(s) - done)
        done = int(s)
        assert np.array_equal(rs.priv_bodies(), g["bodies_%d" % k], equal_nan=True)
        arbs, hi = rs.priv_arbiters()
        assert np.array_equal(arbs, g["arbiters_%d" % k]) and np.array_equal(hi, g["hash_hi_%d" % k])
        assert np.array_equal(rs.shape_bbs(), g["bbs_%d" % k])


def test_golden_scenes_match_the_reference_demo_code(ref):
    for name in golden_names():
        blob, dt = ref.demo_scene(name)
        sc = golden_scene(name)
        assert blob == sc.blob, name
        assert dt == sc.dt


def test_demo_space_and_reloaded_blob_agree_bit_exactly(ref):
    """A space built by the reference's demo code and the same scene re-created from its blob through the
    public API evolve identically: the blob loses nothing (incl. shape hashids / insertion order).
    (Bench.c's boxes are excluded: add_box inserts the shape with radius 0 and bevels it afterwards, so the
    demo's BBTree is built from smaller leaves than a reload's, which permutes the arbiter order.)"""
    import ctypes as C
    for name in ("SimpleTerrainHexagons_100", "Chains"):
        space, dt = ref.demo_space(name)
        blob, _ = ref.demo_scene(name)
        rs = ref.load(blob)
        ref.cp.refp_step(space, dt, 150)
        rs.step(dt, 150)
        n = rs.n_bodies
        a = np.full((n, 24), np.nan)
        ref.cp.refp_get_bodies(space, n, a.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.array_equal(a, rs.priv_bodies(), equal_nan=True)


def test_reference_passes_its_restated_known_answers(ref, tmp_path):
    from oracle.ref import REF_DIR
    inc = "/root/reference/include"
    if not os.path.isdir(inc):
        pytest.skip("reference headers not present")
    exe = str(tmp_path / "known_answers_ref")
    subprocess.check_call(["gcc", "-O1", "-w", "-o", exe, os.path.join(ROOT, "tests/c/known_answers.c"), "-I", inc,
                           "-L", REF_DIR, "-lchipmunk_ref", "-Wl,-rpath," + REF_DIR, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.count("PASS") == 8, out.stdout


def test_edge_cases_golden_is_the_reference_output(ref, tmp_path):
    """tests/golden/edge_cases.ref.txt (what tests/test_gpu_edges.py holds the drop-in to) is the output of
    tests/c/edge_cases.c linked against the unmodified reference -- regenerate it with exactly this recipe."""
    from oracle.ref import REF_DIR
    exe = str(tmp_path / "edge_cases_ref")
    subprocess.check_call(["gcc", "-O1", "-w", "-o", exe, os.path.join(ROOT, "tests/c/edge_cases.c"), "-I", os.path.join(ROOT, "include"),
                           "-L", REF_DIR, "-lchipmunk_ref", "-Wl,-rpath," + REF_DIR, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    with open(os.path.join(ROOT, "tests/golden/edge_cases.ref.txt")) as f:
        assert out.stdout == f.read()


def test_pair_set_is_index_independent(ref):
    """the surviving pair set equals brute force over cached AABBs; spot-check that the
    probe's brute force agrees with the arbiters the reference actually created (every arbiter's pair is in it)."""
    sc = golden_scene("SimpleTerrainCircles_100")
    rs = ref.load(sc.blob)
    rs.step(sc.dt, 120)
    pairs = set(int(p) for p in rs.pairs())
    arbs, _ = rs.priv_arbiters()
    for r in arbs:
        lo, hi = int(min(r[0], r[1])), int(max(r[0], r[1]))
        assert ((lo << 32) | hi) in pairs
