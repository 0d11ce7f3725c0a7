"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_ref(name):
    return np.load(os.path.join(GOLDEN, name + ".ref.npz"))


def order_keys(arbs):
    """space->arbiters rows (oracle) -> the (shape a << 32 | shape b) list cpb200_world_set_arbiter_order takes."""
    return (arbs[:, 0].astype(np.uint64) << np.uint64(32)) | arbs[:, 1].astype(np.uint64)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.nanmax(np.abs(a - b) / (1.0 + np.maximum(np.abs(a), np.abs(b))))) if a.size else 0.0


def match_arbiters(ref_rows, ref_hash_hi, dev):
    """Pair oracle arbiter rows (ref_probe.c REFP_ARB_ROW) with device arbiters by unordered shape pair.
    Yields (ref_row, hash_hi_row, dev_record, swapped)."""
    table = {}
    for d in dev:
        table[(min(int(d["shape_a"]), int(d["shape_b"])), max(int(d["shape_a"]), int(d["shape_b"])))] = d
    for r, hi in zip(ref_rows, ref_hash_hi):
        key = (int(min(r[0], r[1])), int(max(r[0], r[1])))
        d = table.get(key)
        yield r, hi, d, (d is not None and int(r[0]) != int(d["shape_a"]))


def oracle_body_descs(ref_rows, scene, dev_state):
    """refp_get_bodies rows -> cpb200_body_desc records (sleep bookkeeping taken from the device)."""
    from chipmunk2d_b200.engine import BODY_DESC
    n = len(ref_rows)
    d = np.zeros(n, dtype=BODY_DESC)
    r = np.nan_to_num(ref_rows, nan=0.0, posinf=np.inf, neginf=-np.inf)
    d["p"] = r[:, 0:2]; d["v"] = r[:, 2:4]; d["a"] = r[:, 4]; d["w"] = r[:, 5]
    d["v_bias"] = r[:, 6:8]; d["w_bias"] = r[:, 8]; d["f"] = r[:, 9:11]; d["t"] = r[:, 11]
    d["rot"] = r[:, 12:14]
    d["idle_time"] = r[:, 18]
    d["m"] = scene.bodies["m"]; d["i"] = scene.bodies["i"]; d["cog"] = scene.bodies["cog"]
    d["type"] = scene.bodies["type"]
    d["sleeping"] = dev_state["sleeping"]; d["sleep_group"] = dev_state["sleep_group"]
    # row 0 (the space's static body) is not reported by cpSpaceEachBody: keep the scene's values
    d["rot"][0] = (1.0, 0.0); d["idle_time"][0] = np.inf
    return d


def lockstep(ref_space, world, dt, steps, check=None, resync_scene=None):
    """Step the oracle and the device side by side with the device solving in the oracle's order.
    With resync_scene the device's body state is overwritten with the oracle's before every step, so
    each step is a one-step parity check from identical inputs (arbiter warm-start state stays the
    device's own)."""
    for s in range(steps):
        if resync_scene is not None:
            world.update_bodies(0, oracle_body_descs(ref_space.priv_bodies(), resync_scene, world.bodies()))
        asleep = np.nan_to_num(ref_space.priv_bodies()[:, 19]).astype(np.uint8)
        ref_space.step(dt)
        arbs, hi = ref_space.priv_arbiters()
        world.set_arbiter_order(order_keys(arbs))
        if ref_space.n_joints:
            world.set_joint_order(ref_space.constraint_order())
        world.step(dt)
        world.sync()
        if check:
            check(s + 1, asleep, arbs, hi)
