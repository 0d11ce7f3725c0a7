"""Dev tool: step two identical worlds side by side and report the first step where their states differ.
usage: gpu_determinism.py SCENE STEPS [grid_a grid_b]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import golden_scene

name = sys.argv[1]; steps = int(sys.argv[2])
ga = int(sys.argv[3]) if len(sys.argv) > 3 else 0
gb = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sc = golden_scene(name)
A = World(1); A.load_scene(sc)
B = World(1); B.load_scene(sc)
if ga: A.set_solver_grid(ga)
if gb: B.set_solver_grid(gb)
first = None
for i in range(steps):
    A.step(sc.dt); B.step(sc.dt)
    a = A.bodies(); b = B.bodies()
    same = all(np.array_equal(a[k], b[k]) for k in ("p", "v", "a", "w", "sleeping"))
    if not same:
        first = i
        d = np.abs(a["v"] - b["v"]).max(axis=1)
        bad = np.nonzero(d)[0]
        print("first divergence at step", i, "bodies", bad[:10], "max dv", d.max(), "sleeping", a["sleeping"].sum(), b["sleeping"].sum())
        print("stats A", A.stats()); print("stats B", B.stats())
        break
print(name, "grids", ga, gb, "hints", "off" if os.environ.get("CPB200_NO_HINTS") else "on", "->", "DIVERGED at %d" % first if first is not None else "identical for %d steps" % steps)
