import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import mixed_drop, circle_pile
for name, sc, warm in (("mixed100k", mixed_drop(100000), 120), ("pile100k", circle_pile(100000, dense=True, sleep=np.inf), 60)):
    for g in (0, 296, 148, 96, 64, 32, 16):
        w = World(1); w.load_scene(sc)
        if g: w.set_solver_grid(g)
        w.step(sc.dt, warm); w.sync()
        ms = w.time_steps(sc.dt, 100)/100
        w.set_profiling(True)
        for _ in range(3): w.step(sc.dt)
        sp = w.solver_profile(); st = w.stage_times()
        print("%-10s grid %3d  %.3f ms/step  solve %.0f us (colour %.0f rows %.0f warm %.0f iterate %.0f)" % (name, g, ms, st["solve"], sp["colour_us"], sp["rows_us"], sp["warm_us"], sp["iterate_us"]), flush=True)
        w.close()
