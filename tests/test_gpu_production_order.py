"""The solver kernels that SHIP (graph-coloured order) pinned against the CPU oracle.

Row a18/a19 of SURVEY.md 8(a): cpArbiterApplyCachedImpulse / cpArbiterApplyImpulse (cpArbiter.c:441-498) and the joints'
applyCachedImpulse / applyImpulse, iterated by cpSpaceStep.c:406-427.  The production kernels solve colour phase after
colour phase; tests/replay.py reads the solver's inputs right before K10/K11 (validation hooks of include/cpb200.h),
asks the device for the sequence it used, replays that sequence through oracle/cp_oracle.c's sequential solver
(pinned bit for bit against the unmodified reference by tests/test_cpu_oracle_restatement.py) and compares every
output of the solver: v, w, v_bias, w_bias of every body, jnAcc / jtAcc / jBias of every contact, the accumulated
impulse of every joint.

Tolerance: 1e-9 relative is the contract (north star); the kernels are compiled -fmad=false with the reference's
operation order, so the results are expected -- and asserted -- to be BIT-IDENTICAL.

All three kernel families are pinned: the world-wide persistent kernel with L2-cached rows (k_colour_solve<0,0>), the
same with streamed rows (<0,1>), and the space-local one-CTA-per-space solver (k_sl_rows + k_sl_solve<0/1>).
"""
import numpy as np
import pytest

from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import golden_scene, circle_pile, mixed_drop, batched_demo_scenes, all_joints_scene, hub_scene
from tests.replay import production_step_replay

pytestmark = pytest.mark.gpu

TOL = 1e-9
CACHED, STREAMED, SPACE_LOCAL = 1, 2, 3


def scenes_for(name):
    if name == "batch8":
        return batched_demo_scenes(8)
    if name == "all_joints":
        return [all_joints_scene()]
    if name == "hub_300":
        return [hub_scene(300)]
    if name == "mixed_drop_6000":
        return [mixed_drop(6000)]
    if name == "circle_pile_20000":
        return [circle_pile(20000, dense=True, sleep=0.5)]
    return [golden_scene(name)]


# (scene, steps, every-nth step is replayed after the first five)
CASES = [("SimpleTerrainCircles_1000", 50, 1), ("ComplexTerrainHexagons_1000", 50, 1), ("PyramidStack", 50, 1), ("Chains", 50, 1),
         ("all_joints", 50, 1), ("mixed_drop_6000", 50, 5), ("circle_pile_20000", 50, 10), ("batch8", 50, 5)]
# untimed production steps before the replayed ones, so that contacts exist (the pyramid's boxes start apart)
SETTLE = {"PyramidStack": 260, "batch8": 260, "SimpleTerrainCircles_1000": 30, "ComplexTerrainHexagons_1000": 30, "mixed_drop_6000": 20}
SMALL = ("SimpleTerrainCircles_1000", "ComplexTerrainHexagons_1000", "PyramidStack", "Chains", "all_joints", "batch8")


def run_case(name, variant, steps, every, grid=0):
    scenes = scenes_for(name)
    w = World(len(scenes))
    w.load_scenes(scenes)
    w.set_solver_variant(variant)
    if grid:
        w.set_solver_grid(grid)
    dt = scenes[0].dt
    iterations = max(int(sc.header["iterations"]) for sc in scenes)
    worst = 0.0
    items = joints = 0
    if grid == 0:
        first = production_step_replay(w, dt, 0.0, iterations)       # the very first step: dt_coef = 0, nothing cached
        assert first.max_rel <= TOL and first.bit_equal, (name, variant, "first step", first.detail)
        w.step(dt, SETTLE.get(name, 0))
    for s in range(0 if grid else 1, steps):
        if s < 5 or (s + 1) % every == 0 or s == steps - 1:
            r = production_step_replay(w, dt, 0.0 if (grid and s == 0) else 1.0, iterations)
            assert r.path == (2 if variant == SPACE_LOCAL else 1), (name, variant, r.path)
            assert r.max_rel <= TOL, (name, variant, s, r.max_rel, r.detail)
            assert r.bit_equal, (name, variant, s, r.detail)
            worst = max(worst, r.max_rel)
            items += r.n_items; joints += r.n_joints
        else:
            w.step(dt)
    w.sync()
    st = w.stats()
    assert st["overflow"] == 0
    assert items > 0
    return st, joints


@pytest.mark.parametrize("name,steps,every", CASES)
@pytest.mark.parametrize("variant", [CACHED, STREAMED])
def test_world_wide_coloured_kernel_equals_oracle_replay(name, steps, every, variant):
    """k_colour_solve<false, false> (rows through L2) and <false, true> (streamed rows), automatic grid."""
    st, joints = run_case(name, variant, steps, every)
    if name in ("Chains", "all_joints", "mixed_drop_6000", "batch8"):
        assert joints > 0


@pytest.mark.parametrize("name", ["ComplexTerrainHexagons_1000", "Chains", "all_joints"])
@pytest.mark.parametrize("grid", [7, 1000])
def test_world_wide_kernel_on_many_ctas_equals_oracle_replay(name, grid):
    """The same small scenes forced onto 7 CTAs and onto the full co-resident grid: the grid barrier path."""
    run_case(name, STREAMED, 30, 1, grid=grid)


@pytest.mark.parametrize("name,steps,every", [c for c in CASES if c[0] in SMALL])
def test_space_local_kernel_equals_oracle_replay(name, steps, every):
    """k_sl_rows + k_sl_solve<false> / <true>: packed rows, velocities in shared memory."""
    st, joints = run_case(name, SPACE_LOCAL, steps, every)
    if name in ("Chains", "all_joints", "batch8"):
        assert joints > 0


def test_automatic_choice_runs_the_pinned_families():
    """What the engine picks by itself is one of the families above: space-local for the batch, world-wide for a pile."""
    sc = batched_demo_scenes(8)
    w = World(8); w.load_scenes(sc); w.step(sc[0].dt, 3); w.sync()
    assert w.solver_path() == 2
    pile = circle_pile(20000, dense=True)
    w = World(1); w.load_scene(pile); w.step(pile.dt, 3); w.sync()
    assert w.solver_path() == 1


@pytest.mark.parametrize("variant", [STREAMED, SPACE_LOCAL])
def test_body_with_300_contacts_is_fully_solved(variant):
    """A dynamic body of degree 300: 63 of its contacts get regular colours, the rest the serial overflow bucket --
    every one of them is solved (the replay asserts the order covers all active arbiters exactly once) and the result
    equals the sequential replay.  Regression for contacts that were silently left unsolved when the colouring
    rounds ran out."""
    st, _ = run_case("hub_300", variant, 12, 1)
    assert st["n_arbiters"] >= 300
    assert st["n_colours"] == 64          # the overflow bucket is in use
