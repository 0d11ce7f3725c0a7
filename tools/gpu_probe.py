#!/usr/bin/env python
"""Quick on-GPU probe: ms/step and per-stage device times for a few scenes (development aid)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chipmunk2d_b200.engine import World  # noqa: E402
from chipmunk2d_b200.scenes import golden_scene, circle_pile, mixed_drop, batched_demo_scenes  # noqa: E402


def probe(label, scenes, warm, steps):
    sc0 = scenes[0]
    t0 = time.time()
    w = World(len(scenes))
    w.load_scenes(scenes)
    t_load = time.time() - t0
    w.step(sc0.dt, warm); w.sync()
    t0 = time.time()
    w.step(sc0.dt, steps); w.sync()
    dt_ms = (time.time() - t0) * 1000.0 / steps
    st = w.stats()
    w.set_profiling(True)
    acc = {}
    for _ in range(5):
        w.step(sc0.dt)
        for k, v in w.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / 5.0
    w.set_profiling(False)
    nb = sum(s.n_dynamic() for s in scenes)
    print("%-28s bodies %8d load %.1fs  %.3f ms/step  %.3e body-steps/s  arbs %d contacts %d colours %d pairs %d awake %d overflow %d" % (
        label, nb, t_load, dt_ms, nb / (dt_ms * 1e-3), st["n_arbiters"], st["n_contacts"], st["n_colours"], st["n_pairs"], st["n_awake"], st["overflow"]))
    print("    stages(us): " + "  ".join("%s %.0f" % (k, v) for k, v in acc.items()), flush=True)
    print("    solver: " + "  ".join("%s %.0f" % (k, v) for k, v in w.solver_profile().items()), flush=True)
    w.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "pile100k", "pile1m", "pile1m_nosleep", "mixed100k", "batch1024", "batch4096"]
    for name in which:
        if name == "c1":
            probe("SimpleTerrainCircles_1000", [golden_scene("SimpleTerrainCircles_1000")], 100, 500)
        elif name == "c2":
            probe("ComplexTerrainHexagons_1000", [golden_scene("ComplexTerrainHexagons_1000")], 100, 500)
        elif name == "pile100k":
            probe("circle_pile 100k dense", [circle_pile(100000, dense=True)], 60, 60)
        elif name == "pile1m":
            probe("circle_pile 1M dense", [circle_pile(1000000, dense=True)], 40, 40)
        elif name == "pile1m_nosleep":
            probe("circle_pile 1M dense nosleep", [circle_pile(1000000, dense=True, sleep=np.inf)], 40, 40)
        elif name == "mixed100k":
            probe("mixed_drop 100k", [mixed_drop(100000)], 60, 60)
        elif name.startswith("batch"):
            n = int(name[5:])
            probe("batched demos x%d" % n, batched_demo_scenes(n), 300, 200)
