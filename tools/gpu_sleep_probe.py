"""Developer probe: a sleeping 1M pile is hit -- wall time of the steps around the wake, with and without the step graph."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from chipmunk2d_b200.engine import World, scene_descs
from chipmunk2d_b200.scenes import circle_pile

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
prof = len(sys.argv) > 2 and sys.argv[2] == "prof"
sc = circle_pile(n, dense=True, sleep=0.5, columns=max(8, n // 20))
w = World(1); w.load_scene(sc)
t = time.time(); w.step(sc.dt, 700); w.sync(); print("700 steps %.2fs" % (time.time() - t), w.stats()["n_awake"], flush=True)
b1 = w.bodies()
bd, _, _ = scene_descs(sc)
k = int(np.argmax(b1["p"][1:, 1])) + 1
d = bd[k:k + 1].copy()
d["p"] = b1["p"][k]; d["a"] = b1["a"][k]; d["rot"] = b1["rot"][k]; d["v"] = (400.0, -50.0); d["sleeping"] = 0; d["sleep_group"] = -1
t = time.time(); w.update_bodies(k, d); print("update %.3fs" % (time.time() - t), flush=True)
w.set_profiling(prof)
t = time.time(); w.step(sc.dt, 3); w.sync(); print("3 steps after the kick %.3fs" % (time.time() - t), w.graph_stats(), flush=True)
for s in range(6):
    t = time.time(); w.step(sc.dt); w.sync(); el = time.time() - t
    st = w.stats()
    print("step %d %.3fs awake %d arbs %d colours %d" % (s, el, st["n_awake"], st["n_arbiters"], st["n_colours"]),
          {k: round(v) for k, v in w.stage_times().items()} if prof else w.graph_stats(), flush=True)
t = time.time(); a = w.bodies(); print("bodies %.3fs" % (time.time() - t))
t = time.time(); a = w.pairs(); print("pairs %.3fs" % (time.time() - t), len(a))
t = time.time(); a = w.shape_bbs(); print("bbs %.3fs" % (time.time() - t))
