"""Development aid: larger mixed scenes (polygons, joints) -- looks for pathologies outside the circle pile."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chipmunk2d_b200.engine import World
from chipmunk2d_b200.scenes import mixed_drop
for n in (300000, 1000000):
    sc = mixed_drop(n)
    w = World(1); w.load_scene(sc)
    w.step(sc.dt, 150); w.sync()
    ms = w.time_steps(sc.dt, 30)/30
    w.set_profiling(True)
    for _ in range(3): w.step(sc.dt)
    st = w.stats(); sp = w.solver_profile(); tm = w.stage_times()
    print("mixed %d: %.3f ms/step %.3e body-steps/s arbs %d contacts %d colours %d overflow %d maxpen %.2f" % (n, ms, n/(ms*1e-3), st["n_arbiters"], st["n_contacts"], st["n_colours"], st["overflow"], st["max_penetration"]))
    print("   " + "  ".join("%s %.0f" % kv for kv in tm.items()))
    print("   " + "  ".join("%s %.0f" % kv for kv in sp.items()), flush=True)
    w.close()
