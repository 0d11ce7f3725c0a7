#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref = /root/reference compiled by oracle/Makefile, no -ffast-math).

  <Scene>.scene      the cpSpace built by the reference's own demo code (demo/Bench.c bench_list,
                     demo/PyramidStack.c, demo/Chains.c; srand(45073) as RunDemo does), flattened into
                     a scene blob (chipmunk2d_b200/scenes/cpb_scene.h) by oracle/ref_probe.c
  <Scene>.ref.npz    reference results for that scene stepped with cpSpaceStep at the demo's dt:
                       steps        step numbers sampled
                       bodies[k]    refp_get_bodies rows (p v a w v_bias w_bias f t transform idle sleeping ...)
                       bbs[k]       cached shape AABBs
                       pairs[k]     overlapping-pair set (sorted min<<32|max of scene shape indices)
                       arbiters[k]  space->arbiters rows in solver order (+ hash_hi)
                       counts[k]    dynamic bodies / arbiters / contacts / sleeping components

Run where /root/reference exists:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref import Ref  # noqa: E402

SCENES = {
    "SimpleTerrainCircles_1000": [1, 2, 10, 100],
    "SimpleTerrainCircles_100": [1, 2, 10, 100, 300],
    "SimpleTerrainBoxes_100": [1, 2, 10, 100, 300],
    "SimpleTerrainHexagons_100": [1, 2, 10, 100, 300],
    "SimpleTerrainVCircles_200": [1, 10, 100],
    "SimpleTerrainVBoxes_200": [1, 10, 100],
    "ComplexTerrainHexagons_1000": [1, 2, 10, 100],
    "SimpleTerrainBoxes_1000": [1, 10, 100],
    "BouncyTerrainHexagons_500": [1, 10, 100],
    "PyramidStack": [1, 10, 100, 300, 600],
    "Chains": [1, 10, 100, 300],
}


def main():
    ref = Ref()
    for name, steps in SCENES.items():
        blob, dt = ref.demo_scene(name)
        with open(os.path.join(HERE, name + ".scene"), "wb") as f:
            f.write(blob)
        rs = ref.load(blob)
        out = {"steps": np.array(steps), "dt": np.array(dt)}
        done = 0
        for k, s in enumerate(steps):
            # the pair set is evaluated with the sleeping flags as they were at collision time
            rs.step(dt, s - done - 1)
            asleep = rs.priv_bodies()[:, 19]
            asleep = np.nan_to_num(asleep).astype(np.uint8)
            rs.step(dt, 1)
            done = s
            arbs, hi = rs.priv_arbiters()
            out["bodies_%d" % k] = rs.priv_bodies()
            out["bbs_%d" % k] = rs.shape_bbs()
            out["pairs_%d" % k] = rs.pairs(asleep)
            out["arbiters_%d" % k] = arbs
            out["hash_hi_%d" % k] = hi
            out["joint_order_%d" % k] = rs.constraint_order()
            c = rs.counts()
            out["counts_%d" % k] = np.array([c["dynamic_bodies"], c["arbiters"], c["contacts"], c["sleeping_components"]])
        np.savez_compressed(os.path.join(HERE, name + ".ref.npz"), **out)
        print(name, len(blob), "bytes; steps", steps)


if __name__ == "__main__":
    main()
