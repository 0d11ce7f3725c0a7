"""Developer probe: step time and broadphase stage times against the LBVH rebuild period (env CPB200_BVH_PERIOD)."""
import os, subprocess, sys, json
code = r'''
import sys, json, os
import numpy as np
sys.path.insert(0, ".")
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import circle_pile, mixed_drop, golden_scene
name = sys.argv[1]
def explosion():
    """200 k circles thrown about at up to 3 diameters per step: the worst case for a tree whose topology is kept"""
    from chipmunk2d_b200.engine import Scene
    sc = circle_pile(200000, sleep=float("inf"))
    b = sc.bodies.copy()
    rng = np.random.default_rng(3)
    b["v"][1:] = rng.uniform(-1800.0, 1800.0, size=(len(b) - 1, 2))
    hdr = sc.header.copy(); hdr["gravity"] = (0.0, 0.0)
    return Scene.build(hdr, b, sc.shapes, sc.verts, sc.joints)
sc = {"pile1m": lambda: circle_pile(1000000, dense=True, sleep=0.5), "mixed100k": lambda: mixed_drop(100000), "explosion": explosion,
      "c1": lambda: golden_scene("SimpleTerrainCircles_1000"), "c2": lambda: golden_scene("ComplexTerrainHexagons_1000")}[name]()
w = World(1); w.load_scene(sc)
settle = {"pile1m": 25, "mixed100k": 120, "c1": 300, "c2": 300, "explosion": 4}[name]
w.step(sc.dt, settle); w.sync()
n = 48
ms = w.time_steps(sc.dt, n) / n if hasattr(w, "time_steps") else 0
w.set_profiling(True)
acc = {}
for s in range(16):
    w.step(sc.dt); w.sync()
    for k, v in w.stage_times().items(): acc[k] = acc.get(k, 0.0) + v / 16
print(name, os.environ.get("CPB200_BVH_PERIOD"), "no valve" if os.environ.get("CPB200_BVH_NO_VALVE") else "valve", "ms/step %.4f" % ms, {k: round(v) for k, v in acc.items() if k.startswith("bvh") or k == "collide"}, w.stats()["n_pairs"])
'''
for name in sys.argv[1:] or ["pile1m", "mixed100k", "c1", "c2"]:
    for period in (("1", "8", "8nv", "32nv") if name == "explosion" else ("1", "2", "4", "8", "16", "32")):
        env = dict(os.environ, CPB200_BVH_PERIOD=period.replace("nv", ""))
        if period.endswith("nv"):
            env["CPB200_BVH_NO_VALVE"] = "1"
        subprocess.run([sys.executable, "-c", code, name], env=env)
