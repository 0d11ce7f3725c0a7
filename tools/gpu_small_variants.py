"""Developer probe (GPU box): solver families on the 1000-body Bench scenes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import golden_scene
for name in ("SimpleTerrainCircles_1000", "ComplexTerrainHexagons_1000"):
    sc = golden_scene(name)
    for variant, grid in ((0, 0), (1, 0), (2, 0), (1, 4), (1, 16)):
        w = World(1); w.load_scene(sc); w.set_solver_variant(variant)
        if grid: w.set_solver_grid(grid)
        w.step(sc.dt, 300); w.sync()
        ms = min(w.time_steps(sc.dt, 200) for _ in range(3)) / 200
        print(name, "variant", variant, "grid", grid, "path", w.solver_path(), "%.4f ms/step" % ms, w.graph_stats(), flush=True)
