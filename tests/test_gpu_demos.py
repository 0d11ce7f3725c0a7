"""The reference's own demo programs, compiled UNMODIFIED from /root/reference/demo at build time (oracle/Makefile:
demo/Bench.c, PyramidStack.c, Chains.c, Planet.c, Springies.c + oracle/demo_driver.c in place of the GLFW framework),
once against the reference and once against the drop-in's include/chipmunk/chipmunk.h + libchipmunk_b200.so -- the latter
steps on the GPU.  Both executables print every body's state after N of the demo's own update calls.

  Planet.c     custom velocity function (planetGravityVelocityFunc) on every box: while the boxes orbit without touching,
               the drop-in must agree with the reference to 1e-9 (slow path of host/cp_space.c).
  Springies.c  custom spring force function (clamped) on 20 springs, several per body: 1e-6 over the first steps, bounded
               and finite later (another Gauss-Seidel order over the springs of one body).
  Bench.c      all 17 scenes step, stay finite and land where the reference's do (statistics: the coloured order).
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "demo_ref")
B200_EXE = os.path.join(ROOT, "oracle", "_ref", "demo_b200")


def run(exe, name, steps):
    out = subprocess.run([exe, name, str(steps)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return np.loadtxt(out.stdout.splitlines(), ndmin=2)


@pytest.fixture(scope="module")
def exes():
    if not (os.path.exists(REF_EXE) and os.path.exists(B200_EXE)):
        pytest.skip("oracle/_ref/demo_ref / demo_b200 not built (needs /root/reference at build time)")
    return REF_EXE, B200_EXE


def rel(a, b):
    return float(np.max(np.abs(a - b) / (1.0 + np.maximum(np.abs(a), np.abs(b)))))


def test_planet_demo_custom_velocity_function(exes):
    r, d = run(exes[0], "Planet", 120), run(exes[1], "Planet", 120)
    assert r.shape == d.shape and len(r) > 20
    assert rel(r[:, 1:], d[:, 1:]) < 1e-9
    d2 = run(exes[1], "Planet", 900)              # boxes have landed on the planet by now: still sane
    assert np.all(np.isfinite(d2)) and np.max(np.hypot(d2[:, 1], d2[:, 2])) < 2000.0


def test_springies_demo_custom_spring_force_function(exes):
    r, d = run(exes[0], "Springies", 3), run(exes[1], "Springies", 3)
    assert r.shape == d.shape and len(r) > 10
    assert rel(r[:, 1:], d[:, 1:]) < 1e-6
    r2, d2 = run(exes[0], "Springies", 300), run(exes[1], "Springies", 300)
    assert np.all(np.isfinite(d2))
    assert np.max(np.abs(d2[:, 1:3])) < 2.0 * np.max(np.abs(r2[:, 1:3])) + 100.0


@pytest.mark.parametrize("name", ["SimpleTerrainCircles_1000", "SimpleTerrainBoxes_500", "SimpleTerrainVHexagons_200",
                                  "ComplexTerrainHexagons_1000", "BouncyTerrainCircles_500", "NoCollide", "PyramidStack", "Chains"])
def test_bench_and_stacking_demos_run_unmodified_on_the_gpu(exes, name):
    steps = 150
    r, d = run(exes[0], name, steps), run(exes[1], name, steps)
    assert r.shape == d.shape
    assert np.all(np.isfinite(d))
    # the pile sits where the reference's does (different Gauss-Seidel order: statistics, not bits)
    assert abs(np.mean(d[:, 2]) - np.mean(r[:, 2])) < 5.0 + 0.02 * abs(np.mean(r[:, 2]))
    assert np.min(d[:, 2]) > np.min(r[:, 2]) - 10.0
    if name == "NoCollide":
        assert rel(r[:, 1:], d[:, 1:]) < 1e-9
