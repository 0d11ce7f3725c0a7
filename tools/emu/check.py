#!/usr/bin/env python
"""Developer tool (GPU-less container): run the CPB_EMU build of the kernels against the
reference oracle.  Not part of the product or of tests/.  Usage: tools/emu/check.py [scene] [steps] [mode]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from chipmunk2d_b200.engine import World, Scene  # noqa: E402
from oracle.ref import Ref  # noqa: E402

EMU = os.path.join(ROOT, "tools/emu/_build/libcpb200_emu.so")


def compare(name, steps, mode):
    ref = Ref()
    blob, dt = ref.demo_scene(name)
    sc = Scene(blob)
    rs = ref.load(blob)
    w = World(1, lib_path=EMU)
    w.load_scene(sc)
    w.set_solver_mode(mode)
    nb = sc.header["n_bodies"]
    print(name, "bodies", nb, "shapes", len(sc.shapes), "joints", len(sc.joints), "dt", dt)
    bb_ref = rs.shape_bbs(); bb = w.shape_bbs()
    print("  initial bb equal:", np.array_equal(bb_ref, bb))
    for s in range(steps):
        rs.step(dt)
        arbs, hi = rs.priv_arbiters()
        if mode == 1:
            order = (arbs[:, 0].astype(np.uint64) << np.uint64(32)) | arbs[:, 1].astype(np.uint64)
            w.set_arbiter_order(order)
            w.set_joint_order(rs.constraint_order())
        w.step(dt)
        w.sync()
        st = w.stats()
        pr = rs.pairs()
        pw = w.pairs()
        rb = rs.priv_bodies()
        wb = w.bodies()
        dp = np.nanmax(np.abs(rb[:, 0:2] - wb["p"])); dv = np.nanmax(np.abs(rb[:, 2:4] - wb["v"]))
        dw = np.nanmax(np.abs(rb[:, 5] - wb["w"]))
        wa = w.arbiters()
        if (s + 1) % max(1, steps // 6) == 0 or not np.array_equal(pr, pw): print("  step %d pairs ref %d dev %d equal %s | arbs ref %d dev %d contacts %d colours %d | max|dp| %.3g |dv| %.3g |dw| %.3g" % (
            s + 1, len(pr), len(pw), np.array_equal(pr, pw), len(arbs), len(wa), st["n_contacts"], st["n_colours"], dp, dv, dw))
        if st["overflow"]:
            print("  OVERFLOW", st["overflow"])
    return rs, w


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "SimpleTerrainCircles_100"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    compare(name, steps, mode)
