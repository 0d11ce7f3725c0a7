"""The captured step graph (one cudaGraphLaunch per step) against the kernel-by-kernel launch sequence: bit-identical
state, on small latency-bound scenes, on a batch of spaces (two-stream fork inside the graph) and on a pile with the
islands pass; uploads, dt changes and read-backs in the middle of a run invalidate or bypass the graph correctly."""
import numpy as np
import pytest

from chipmunk2d_b200.engine import World, scene_descs
from chipmunk2d_b200.scenes import golden_scene, circle_pile, batched_demo_scenes, mixed_drop

pytestmark = pytest.mark.gpu


def pair(scenes):
    a, b = World(len(scenes)), World(len(scenes))
    a.load_scenes(scenes); b.load_scenes(scenes)
    b.set_graph(False)
    return a, b


def same(a, b):
    x, y = a.bodies(), b.bodies()
    return all(np.array_equal(x[k], y[k]) for k in ("p", "v", "a", "w", "sleeping", "idle_time"))


@pytest.mark.parametrize("name", ["SimpleTerrainCircles_1000", "ComplexTerrainHexagons_1000", "PyramidStack", "Chains", "batch8", "pile20k", "mixed6k"])
def test_graph_replay_equals_kernel_by_kernel_launches(name):
    scenes = {"batch8": lambda: batched_demo_scenes(8), "pile20k": lambda: [circle_pile(20000, dense=True, sleep=0.5)],
              "mixed6k": lambda: [mixed_drop(6000)]}.get(name, lambda: [golden_scene(name)])()
    a, b = pair(scenes)
    dt = scenes[0].dt
    for chunk in range(6):
        a.step(dt, 50); b.step(dt, 50)
        a.sync(); b.sync()
        assert same(a, b), (name, chunk)
    ga, gb = a.graph_stats(), b.graph_stats()
    assert ga["replays"] >= 250 and ga["captures"] >= 2, ga      # one graph per arbiter-buffer parity, then replays
    assert gb["replays"] == 0 and gb["captures"] == 0, gb
    assert a.stats()["overflow"] == 0
    sa, sb = a.stats(), b.stats()
    assert sa["n_arbiters"] == sb["n_arbiters"] and sa["n_contacts"] == sb["n_contacts"]


def test_graph_survives_dt_changes_uploads_and_split_steps():
    sc = golden_scene("ComplexTerrainHexagons_1000")
    a, b = pair([sc])
    dt = sc.dt
    bd, _, _ = scene_descs(sc)
    for w in (a, b):
        w.step(dt, 40)
        w.step(dt * 0.5, 7)                   # dt change: dt_coef != 1 for one step, new pow() terms
        w.step(dt, 25)
        kick = bd[5:6].copy(); st = w.bodies()
        kick["p"] = st["p"][5]; kick["a"] = st["a"][5]; kick["rot"] = st["rot"][5]; kick["v"] = (0.0, 250.0)
        w.update_bodies(5, kick)              # an upload between two steps
        w.step(dt, 30)
        w.step_collide(dt); w.step_finish()   # a split step in between (collision-handler path)
        w.step(dt, 30)
        w.sync()
    assert same(a, b)
    assert a.graph_stats()["replays"] > 100


@pytest.mark.parametrize("name", ["ComplexTerrainHexagons_1000", "pile20k", "batch8"])
def test_bound_io_equals_blocking_uploads_and_downloads(name):
    """cpb200_world_bind_io (forces in / state out on a side stream, overlapped with the step, graph-captured) against
    cpb200_world_set_body_forces + cpb200_world_get_bodies around every step: identical state, every step."""
    scenes = {"batch8": lambda: batched_demo_scenes(8), "pile20k": lambda: [circle_pile(20000, dense=True, sleep=0.5)]}.get(name, lambda: [golden_scene(name)])()
    a, b = World(len(scenes)), World(len(scenes))
    a.load_scenes(scenes); b.load_scenes(scenes)
    n = a.n_bodies
    forces = a.pinned_array(3 * n, np.float64).reshape(n, 3)
    sink = a.pinned_array(6 * n, np.float64).reshape(2, n, 3)
    a.bind_io(forces, sink)
    rng = np.random.default_rng(5)
    dt = scenes[0].dt
    for s in range(120):
        f = rng.normal(size=(n, 3)) * 30.0
        forces[:] = f
        a.step(dt); a.sync()
        b.set_body_forces(0, f); b.step(dt); b.sync()
        if s % 7 == 0 or s > 110:
            wb = b.bodies()
            assert np.array_equal(sink[0][:, 0:2], wb["p"]) and np.array_equal(sink[0][:, 2], wb["a"]), (name, s)
            assert np.array_equal(sink[1][:, 0:2], wb["v"]) and np.array_equal(sink[1][:, 2], wb["w"]), (name, s)
    assert a.graph_stats()["replays"] > 60, a.graph_stats()
    a.bind_io(None, None)
    a.step(dt, 3); b.step(dt, 3); a.sync(); b.sync()
    assert np.array_equal(a.bodies()["p"], b.bodies()["p"])
    with pytest.raises(Exception):
        a.bind_io(np.zeros((n, 3)), None)          # pageable memory is refused


@pytest.mark.parametrize("name", ["ComplexTerrainHexagons_1000", "PyramidStack", "Chains", "mixed6k", "pile20k", "batch8"])
def test_production_step_equals_split_step(name):
    """A production step folds cpArbiterPreStep into the solver's row build and integrates the velocities after it
    (cpArbiterPreStep reads the velocities from before cpBodyUpdateVelocity); a split step -- collision handlers, the
    validation hooks of the oracle replay -- runs the stand-alone kernels in the reference's order.  Same arithmetic on
    the same inputs: identical bits, in the bodies and in the arbiters (nMass / tMass / bias are recomputed on read-back
    after a production step)."""
    scenes = {"batch8": lambda: batched_demo_scenes(8), "pile20k": lambda: [circle_pile(20000, dense=True, sleep=0.5)],
              "mixed6k": lambda: [mixed_drop(6000)]}.get(name, lambda: [golden_scene(name)])()
    a, b = World(len(scenes)), World(len(scenes))
    a.load_scenes(scenes); b.load_scenes(scenes)
    dt = scenes[0].dt
    steps = 320 if name in ("PyramidStack", "batch8") else 80
    for s in range(steps):
        a.step(dt)
        b.step_collide(dt); b.step_presolve(); b.step_finish()
        if s % 20 == 19 or s == steps - 1:
            a.sync(); b.sync()
            assert same(a, b), (name, s)
            ra, rb = a.arbiters(active_only=True), b.arbiters(active_only=True)
            assert len(ra) == len(rb)
            oa, ob = np.argsort(ra["shape_a"].astype(np.int64) << 32 | ra["shape_b"]), np.argsort(rb["shape_a"].astype(np.int64) << 32 | rb["shape_b"])
            ra, rb = ra[oa], rb[ob]
            for f in ("shape_a", "shape_b", "count", "n", "e", "u"):
                assert np.array_equal(ra[f], rb[f]), (name, s, f)
            for k in range(2):
                m = ra["count"] > k
                for f in ("r1", "r2", "n_mass", "t_mass", "bias", "jn_acc", "jt_acc", "j_bias"):
                    assert np.array_equal(ra["contacts"][f][m, k], rb["contacts"][f][m, k]), (name, s, k, f)
    assert a.stats()["n_arbiters"] > 0


def test_graph_is_recaptured_after_structural_edits_reallocate_device_tables():
    """Appending / removing objects re-allocates the space-local tables (and, beyond the slack, the object arrays).  A
    freed array may or may not get its old address back, so equal pointers prove nothing: the graph signature carries
    the allocation generation of every group of device arrays, and a graph captured before the edit is never replayed
    after it.  Other worlds are created in between to stir the device allocator."""
    from tests.test_gpu_append import newcomers
    sc = golden_scene("PyramidStack")
    a, b = pair([sc])
    dt = sc.dt
    a.step(dt, 40); b.step(dt, 40)
    bd, sd, verts, _jd = newcomers(a, 2, "circle", y0=300.0)
    junk = []
    for w in (a, b):
        w.append_bodies(bd); w.append_shapes(sd, verts)
        junk.append(World(1)); junk[-1].load_scene(golden_scene("Chains"))
        w.step(dt, 40)
        w.remove_shape(w.n_shapes - 1)                   # the shape count of two edits ago: same sizes, new tables
        junk.append(World(1)); junk[-1].load_scene(sc)
        w.step(dt, 40)
        w.remove_shape(w.n_shapes - 1)
        w.step(dt, 40)
    a.sync(); b.sync()
    assert same(a, b)
    assert a.stats()["overflow"] == 0 and a.graph_stats()["replays"] > 60, a.graph_stats()
