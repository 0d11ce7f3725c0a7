"""Developer probe: step time and broadphase stage times against the LBVH rebuild period (env CPB200_BVH_PERIOD)."""
import os, subprocess, sys, json
code = r'''
import sys, json, os
import numpy as np
sys.path.insert(0, ".")
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import circle_pile, mixed_drop, golden_scene
name = sys.argv[1]
sc = {"pile1m": lambda: circle_pile(1000000, dense=True, sleep=0.5), "mixed100k": lambda: mixed_drop(100000),
      "c1": lambda: golden_scene("SimpleTerrainCircles_1000"), "c2": lambda: golden_scene("ComplexTerrainHexagons_1000")}[name]()
w = World(1); w.load_scene(sc)
settle = {"pile1m": 25, "mixed100k": 120, "c1": 300, "c2": 300}[name]
w.step(sc.dt, settle); w.sync()
n = 48
ms = w.time_steps(sc.dt, n) / n if hasattr(w, "time_steps") else 0
w.set_profiling(True)
acc = {}
for s in range(16):
    w.step(sc.dt); w.sync()
    for k, v in w.stage_times().items(): acc[k] = acc.get(k, 0.0) + v / 16
print(name, os.environ.get("CPB200_BVH_PERIOD"), "ms/step %.4f" % ms, {k: round(v) for k, v in acc.items() if k.startswith("bvh") or k == "collide"}, w.stats()["n_pairs"])
'''
for name in sys.argv[1:] or ["pile1m", "mixed100k", "c1", "c2"]:
    for period in ("1", "2", "4", "8", "16", "32"):
        env = dict(os.environ, CPB200_BVH_PERIOD=period)
        subprocess.run([sys.executable, "-c", code, name], env=env)
