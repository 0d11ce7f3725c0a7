#!/usr/bin/env python
"""Development aid: time the same scene with differently compiled engine libraries."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import circle_pile, batched_demo_scenes, golden_scene, mixed_drop

def run(lib, scenes, warm, steps):
    w = World(len(scenes), lib_path=lib)
    w.load_scenes(scenes)
    dt = scenes[0].dt
    w.step(dt, warm); w.sync()
    ms = w.time_steps(dt, steps) / steps
    w.set_profiling(True)
    for _ in range(3): w.step(dt)
    sp = w.solver_profile(); st = w.stage_times()
    sp["n_colours"] = w.stats()["n_colours"]
    w.close()
    return ms, sp, st

if __name__ == "__main__":
    libs = sys.argv[1:]
    sets = {"pile1m": ([circle_pile(1000000, dense=True, sleep=np.inf)], 40, 30), "batch4096": (batched_demo_scenes(4096), 300, 100),
            "c1": ([golden_scene("SimpleTerrainCircles_1000")], 200, 300), "c2": ([golden_scene("ComplexTerrainHexagons_1000")], 200, 300),
            "mixed100k": ([mixed_drop(100000)], 120, 60)}
    only = os.environ.get("SETS")
    for name, (sc, warm, steps) in sets.items():
        if only and name not in only.split(","): continue
        for lib in libs:
            ms, sp, st = run(os.path.join(ROOT, "chipmunk2d_b200/lib", lib), sc, warm, steps)
            print("%-10s %-22s %.3f ms/step  solve %.0f us (colour %.0f rows %.0f warm %.0f iterate %.0f) colours %d" % (name, lib, ms, st["solve"], sp["colour_us"], sp["rows_us"], sp["warm_us"], sp["iterate_us"], sp["n_colours"]), flush=True)
            if os.environ.get("STAGES"): print("      " + "  ".join("%s %.0f" % (k, v) for k, v in st.items()), flush=True)
