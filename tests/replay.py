"""Replay of the device's PRODUCTION solver order through the CPU oracle (oracle/cp_oracle.c).

The graph-coloured solver visits its constraints colour phase by colour phase; constraints of one phase share
no dynamic body, so a sequential Gauss-Seidel sweep over the same sequence (cpSpaceStep.c:406-427 with
arbiters and joints merged into one sequence, `cpo_solve_sequence`) must reproduce the parallel result bit for
bit.  One step is split with the validation hooks of include/cpb200.h:

    step_collide -> step_presolve -> [snapshot the solver's inputs] -> step_finish -> [order, outputs]

and the snapshot is fed to the oracle in the order the device reports.  Test infrastructure only.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# numpy mirrors of cp_oracle.h (align=True reproduces the C layout)
CPO_BODY = np.dtype([("p", "<f8", 2), ("v", "<f8", 2), ("v_bias", "<f8", 2), ("cog", "<f8", 2), ("f", "<f8", 2), ("a", "<f8"), ("w", "<f8"),
                     ("w_bias", "<f8"), ("t", "<f8"), ("m_inv", "<f8"), ("i_inv", "<f8"), ("type", "<i4")], align=True)
CPO_CONTACT = np.dtype([("r1", "<f8", 2), ("r2", "<f8", 2), ("nMass", "<f8"), ("tMass", "<f8"), ("bounce", "<f8"), ("jnAcc", "<f8"),
                        ("jtAcc", "<f8"), ("jBias", "<f8"), ("bias", "<f8")], align=True)
CPO_ARBITER = np.dtype([("body_a", "<i4"), ("body_b", "<i4"), ("count", "<i4"), ("first_collision", "<i4"), ("n", "<f8", 2),
                        ("surface_vr", "<f8", 2), ("e", "<f8"), ("u", "<f8"), ("contacts", CPO_CONTACT, 2)], align=True)
CPO_JOINT = np.dtype([("type", "<i4"), ("a", "<i4"), ("b", "<i4"), ("maxForce", "<f8"), ("errorBias", "<f8"), ("maxBias", "<f8"),
                      ("anchorA", "<f8", 2), ("anchorB", "<f8", 2), ("prm", "<f8", 4), ("r1", "<f8", 2), ("r2", "<f8", 2), ("n", "<f8", 2),
                      ("bias2", "<f8", 2), ("jAcc2", "<f8", 2), ("nMass", "<f8"), ("bias", "<f8"), ("jnAcc", "<f8"), ("k", "<f8", 4),
                      ("target_vrn", "<f8"), ("v_coef", "<f8"), ("iSum", "<f8"), ("clamp", "<f8")], align=True)

ARB_FIRST_COLLISION = 0

_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "libcp_oracle.so")
        lib = C.CDLL(path)
        lib.cpo_solve_sequence.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
        lib.cpo_sizeof.restype = C.c_int
        lib.cpo_sizeof.argtypes = [C.c_int]
        assert lib.cpo_sizeof(0) == CPO_BODY.itemsize and lib.cpo_sizeof(1) == CPO_ARBITER.itemsize and lib.cpo_sizeof(2) == CPO_JOINT.itemsize
        _lib = lib
    return _lib


def oracle_bodies(bs):
    """cpb200_world_get_body_solver_state rows -> cpo_body records."""
    b = np.zeros(len(bs), dtype=CPO_BODY)
    b["v"] = bs[:, 0:2]; b["w"] = bs[:, 2]; b["m_inv"] = bs[:, 3]
    b["v_bias"] = bs[:, 4:6]; b["w_bias"] = bs[:, 6]; b["i_inv"] = bs[:, 7]
    return b


def oracle_arbiters(dev):
    """cpb200_arbiter records (after prestep) -> cpo_arbiter records, same order."""
    a = np.zeros(len(dev), dtype=CPO_ARBITER)
    a["body_a"] = dev["body_a"]; a["body_b"] = dev["body_b"]; a["count"] = dev["count"]
    a["first_collision"] = (dev["state"] == ARB_FIRST_COLLISION)
    a["n"] = dev["n"]; a["surface_vr"] = dev["surface_vr"]; a["e"] = dev["e"]; a["u"] = dev["u"]
    for k in range(2):
        c = dev["contacts"][:, k]
        o = a["contacts"][:, k]
        o["r1"] = c["r1"]; o["r2"] = c["r2"]; o["nMass"] = c["n_mass"]; o["tMass"] = c["t_mass"]; o["bounce"] = c["bounce"]
        o["jnAcc"] = c["jn_acc"]; o["jtAcc"] = c["jt_acc"]; o["jBias"] = c["j_bias"]; o["bias"] = c["bias"]
    return a


def oracle_joints(js):
    """cpb200_world_get_joint_solver_state rows (after prestep) -> cpo_joint records with their solver state."""
    j = np.zeros(len(js), dtype=CPO_JOINT)
    if len(js) == 0:
        return j
    j["type"] = js[:, 0].astype(np.int32); j["a"] = js[:, 1].astype(np.int32); j["b"] = js[:, 2].astype(np.int32)
    j["maxForce"] = js[:, 4]; j["maxBias"] = js[:, 5]
    j["r1"] = js[:, 6:8]; j["r2"] = js[:, 8:10]; j["n"] = js[:, 10:12]
    # the device keeps nMass | iSum | clamp in one slot and scalar | vector bias / impulse in one pair
    j["nMass"] = js[:, 12]; j["iSum"] = js[:, 12]; j["clamp"] = js[:, 12]
    j["k"] = js[:, 13:17]
    j["bias"] = js[:, 17]; j["bias2"] = js[:, 17:19]
    j["jnAcc"] = js[:, 19]; j["jAcc2"] = js[:, 19:21]
    j["target_vrn"] = js[:, 21]; j["v_coef"] = js[:, 22]
    j["prm"] = js[:, 23:27]
    return j


def joint_impulses(j):
    """(n, 2) accumulated impulse of cpo_joint records in the device's (acc.x, acc.y) convention."""
    vec = np.isin(j["type"], (2, 3))
    out = np.zeros((len(j), 2))
    out[:, 0] = np.where(vec, j["jAcc2"][:, 0], j["jnAcc"])
    out[:, 1] = np.where(vec, j["jAcc2"][:, 1], 0.0)
    return out


class StepReplay:
    """Result of one production step replayed on the CPU."""

    def __init__(self):
        self.n_items = self.n_arbiters = self.n_joints = 0
        self.bit_equal = True
        self.max_rel = 0.0
        self.path = 0
        self.detail = ""


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.nanmax(np.abs(a - b) / (1.0 + np.maximum(np.abs(a), np.abs(b)))))


def production_step_replay(w, dt, dt_coef, iterations):
    """One step of world `w` in production order, replayed through the oracle.  Returns a StepReplay."""
    lib = oracle_lib()
    w.step_collide(dt)
    w.step_presolve()
    bs0 = w.body_solver_state()
    arbs0 = w.arbiters(active_only=True)
    js0 = w.joint_solver_state()
    w.step_finish()
    w.sync()
    order = w.solver_order()
    bs1 = w.body_solver_state()
    arbs1 = w.arbiters(active_only=True)
    js1 = w.joint_solver_state()

    out = StepReplay()
    out.path = w.solver_path()
    out.n_items = len(order); out.n_arbiters = len(arbs0); out.n_joints = int(np.count_nonzero(js0[:, 3])) if len(js0) else 0
    # record index -> position in the downloaded list
    by_record = np.argsort(arbs0["record"], kind="stable")
    items = order.astype(np.int64).copy()
    isarb = order >= 0
    if np.any(isarb):
        items[isarb] = by_record[np.searchsorted(arbs0["record"][by_record], order[isarb])]
        assert np.array_equal(arbs0["record"][items[isarb]], order[isarb])
    # every active arbiter and every live joint is visited exactly once
    seen_a = np.sort(items[items >= 0]); seen_j = np.sort(-(items[items < 0] + 1))
    assert np.array_equal(seen_a, np.arange(len(arbs0))), "solver order does not cover the active arbiters exactly once"
    live = np.nonzero(js0[:, 3])[0] if len(js0) else np.zeros(0, dtype=np.int64)
    assert np.array_equal(seen_j, live), "solver order does not cover the live joints exactly once"
    assert np.array_equal(arbs0["record"], arbs1["record"])

    B = oracle_bodies(bs0); A = oracle_arbiters(arbs0); J = oracle_joints(js0)
    lib.cpo_solve_sequence(len(items), items.ctypes.data, A.ctypes.data, J.ctypes.data, B.ctypes.data, int(iterations), float(dt), float(dt_coef))

    got = [bs1[:, 0:3], bs1[:, 4:7]]
    want = [np.column_stack([B["v"], B["w"]]), np.column_stack([B["v_bias"], B["w_bias"]])]
    names = ["v,w", "v_bias,w_bias"]
    for k in range(2):
        mask = (arbs0["count"] > k)
        c1 = arbs1["contacts"][:, k]; c0 = A["contacts"][:, k]
        got.append(np.column_stack([c1["jn_acc"], c1["jt_acc"], c1["j_bias"]])[mask])
        want.append(np.column_stack([c0["jnAcc"], c0["jtAcc"], c0["jBias"]])[mask])
        names.append("contact %d jn,jt,jb" % k)
    if len(live):
        got.append(js1[live][:, 19:21]); want.append(joint_impulses(J)[live]); names.append("joint impulses")
        got.append(js1[live][:, 21]); want.append(J["target_vrn"][live]); names.append("spring target_vrn")
    for g, x, name in zip(got, want, names):
        g = np.ascontiguousarray(g, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
        e = rel(g, x)
        out.max_rel = max(out.max_rel, e)
        same = np.array_equal(g.view(np.uint64), x.view(np.uint64)) or np.array_equal(g + 0.0, x + 0.0)
        if not same:
            out.bit_equal = False
            bad = np.nonzero((g + 0.0 != x + 0.0).reshape(len(g), -1).any(axis=1))[0]
            out.detail += "%s: %d rows differ, max rel %.3g; " % (name, len(bad), e)
    return out
