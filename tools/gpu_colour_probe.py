"""Developer probe (GPU box): colour histogram of the solver at a given pile size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import circle_pile, mixed_drop
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sc = circle_pile(n, dense=True, sleep=np.inf) if (len(sys.argv) < 3 or sys.argv[2] == "pile") else mixed_drop(n)
w = World(1); w.load_scene(sc)
for k in (1, 5, 25, 30):
    w.step(sc.dt, k); w.sync()
    rows, joints = w.colour_sizes()
    print("after", w.stats()["steps"], "steps: rows/colour", rows[rows > 0].tolist(), "joints/colour", joints[joints > 0].tolist(), flush=True)
