"""Pin the plain-C restatement (oracle/cp_oracle.c) against the unmodified reference (oracle/_ref):
data captured from reference runs is replayed through the restated stage functions and must come out
BIT-IDENTICAL (same compiler, -ffp-contract=off).  No GPU involved."""
import ctypes as C
import os

import numpy as np
import pytest

from chipmunk2d_b200.engine import Scene, SCENE_HEADER, SCENE_BODY, SCENE_SHAPE, SCENE_JOINT
from chipmunk2d_b200.scenes import golden_scene, all_joints_scene, ERROR_BIAS_DEFAULT, COLLISION_BIAS_DEFAULT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Vec(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double)]


class Shape(C.Structure):
    _fields_ = [("type", C.c_int), ("count", C.c_int), ("hashid", C.c_uint64), ("a", Vec), ("b", Vec), ("n", Vec), ("r", C.c_double),
                ("bb", C.c_double * 4), ("planes", C.POINTER(C.c_double)), ("rot", Vec), ("a_tangent", Vec), ("b_tangent", Vec)]


class Manifold(C.Structure):
    _fields_ = [("count", C.c_int), ("n", Vec), ("p1", Vec * 2), ("p2", Vec * 2), ("hash", C.c_uint64 * 2), ("id", C.c_uint32)]


class Body(C.Structure):
    _fields_ = [("p", Vec), ("v", Vec), ("v_bias", Vec), ("cog", Vec), ("f", Vec), ("a", C.c_double), ("w", C.c_double), ("w_bias", C.c_double),
                ("t", C.c_double), ("m_inv", C.c_double), ("i_inv", C.c_double), ("type", C.c_int)]


class Contact(C.Structure):
    _fields_ = [("r1", Vec), ("r2", Vec), ("nMass", C.c_double), ("tMass", C.c_double), ("bounce", C.c_double), ("jnAcc", C.c_double),
                ("jtAcc", C.c_double), ("jBias", C.c_double), ("bias", C.c_double)]


class Arbiter(C.Structure):
    _fields_ = [("body_a", C.c_int), ("body_b", C.c_int), ("count", C.c_int), ("first_collision", C.c_int), ("n", Vec), ("surface_vr", Vec),
                ("e", C.c_double), ("u", C.c_double), ("contacts", Contact * 2)]


class Joint(C.Structure):
    _fields_ = [("type", C.c_int), ("a", C.c_int), ("b", C.c_int), ("maxForce", C.c_double), ("errorBias", C.c_double), ("maxBias", C.c_double),
                ("anchorA", Vec), ("anchorB", Vec), ("prm", C.c_double * 4), ("r1", Vec), ("r2", Vec), ("n", Vec), ("bias2", Vec), ("jAcc2", Vec),
                ("nMass", C.c_double), ("bias", C.c_double), ("jnAcc", C.c_double), ("k", C.c_double * 4), ("target_vrn", C.c_double),
                ("v_coef", C.c_double), ("iSum", C.c_double), ("clamp", C.c_double)]


@pytest.fixture(scope="module")
def cpo():
    path = os.path.join(ROOT, "oracle", "libcp_oracle.so")
    if not os.path.exists(path):
        pytest.skip("oracle/libcp_oracle.so not built (make -C oracle oracle)")
    lib = C.CDLL(path)
    lib.cpo_collide.argtypes = [C.POINTER(Shape), C.POINTER(Shape), C.POINTER(Manifold)]
    lib.cpo_body_update_position.argtypes = [C.POINTER(Body), C.c_double, C.POINTER(C.c_double)]
    lib.cpo_body_update_velocity.argtypes = [C.POINTER(Body), Vec, C.c_double, C.c_double]
    lib.cpo_arbiter_prestep.argtypes = [C.POINTER(Arbiter), C.POINTER(Body), C.c_double, C.c_double, C.c_double]
    lib.cpo_joint_prestep.argtypes = [C.POINTER(Joint), C.POINTER(Body), C.POINTER(C.c_double), C.c_double]
    lib.cpo_solve.argtypes = [C.c_int, C.POINTER(Arbiter), C.c_int, C.POINTER(Joint), C.POINTER(Body), C.c_int, C.c_double, C.c_double]
    lib.cpo_solve_sequence.argtypes = [C.c_long, C.POINTER(C.c_int64), C.POINTER(Arbiter), C.POINTER(Joint), C.POINTER(Body), C.c_int, C.c_double, C.c_double]
    lib.cpo_pairs.restype = C.c_long
    lib.cpo_pairs.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
    return lib


def make_shape(sc, shapes_priv, planes, bodies_priv, i):
    s = sc.shapes[i]
    out = Shape()
    out.type = int(s["type"]); out.hashid = i; out.r = float(s["r"])
    row = shapes_priv[i]
    out.bb = (C.c_double * 4)(*row[0:4])
    b = bodies_priv[int(s["body"])]
    rot = (b[12], b[13]) if np.isfinite(b[12]) else (1.0, 0.0)
    out.rot = Vec(*rot)
    out.a_tangent = Vec(*s["a_tangent"]); out.b_tangent = Vec(*s["b_tangent"])
    keep = None
    if out.type == 0:
        out.a = Vec(row[4], row[5])
    elif out.type == 1:
        out.a = Vec(row[4], row[5]); out.b = Vec(row[6], row[7]); out.n = Vec(row[8], row[9])
    else:
        out.count = int(s["n_verts"])
        keep = np.ascontiguousarray(planes[int(s["vert_offset"]):int(s["vert_offset"]) + out.count]).copy()
        out.planes = keep.ctypes.data_as(C.POINTER(C.c_double))
    return out, keep


@pytest.mark.parametrize("name,steps", [("ComplexTerrainHexagons_1000", 60), ("SimpleTerrainBoxes_100", 150), ("SimpleTerrainVBoxes_200", 80),
                                         ("PyramidStack", 260), ("Chains", 140), ("SimpleTerrainCircles_100", 100)])
def test_narrowphase_restatement_is_bit_identical(ref, cpo, name, steps):
    sc = golden_scene(name)
    rs = ref.load(sc.blob)
    rs.step(sc.dt, steps)
    shapes_priv = rs.priv_shapes()
    planes = rs.poly_planes(sc.shapes["vert_offset"])
    bodies_priv = rs.priv_bodies()
    pairs = rs.pairs()
    assert len(pairs) > 0
    hits = 0
    for key in pairs[:800]:
        i, j = int(key >> np.uint64(32)), int(key & np.uint64(0xFFFFFFFF))
        a, ka = make_shape(sc, shapes_priv, planes, bodies_priv, i)
        b, kb = make_shape(sc, shapes_priv, planes, bodies_priv, j)
        swapped = a.type > b.type
        if swapped:
            a, b = b, a
        m = Manifold()
        cpo.cpo_collide(C.byref(a), C.byref(b), C.byref(m))
        n_ref, out = rs.shapes_collide(i, j)
        assert m.count == n_ref, (i, j)
        if n_ref == 0:
            continue
        hits += 1
        nx, ny = (-m.n.x, -m.n.y) if swapped else (m.n.x, m.n.y)
        assert (nx, ny) == (out[1], out[2]) or (nx + 0.0, ny + 0.0) == (out[1] + 0.0, out[2] + 0.0)
        for k in range(n_ref):
            p1, p2 = m.p1[k], m.p2[k]
            pa, pb = ((p2, p1) if swapped else (p1, p2))
            assert [pa.x, pa.y, pb.x, pb.y] == list(out[3 + 5 * k: 7 + 5 * k]), (i, j, k)
    assert hits > 0


def bodies_from_priv(rows, sc):
    n = len(rows)
    arr = (Body * n)()
    for i in range(n):
        r = np.nan_to_num(rows[i], nan=0.0, posinf=np.inf)
        b = arr[i]
        b.p = Vec(r[0], r[1]); b.v = Vec(r[2], r[3]); b.a = r[4]; b.w = r[5]
        b.v_bias = Vec(r[6], r[7]); b.w_bias = r[8]; b.f = Vec(r[9], r[10]); b.t = r[11]
        b.m_inv = r[20]; b.i_inv = r[21]; b.cog = Vec(r[22], r[23])
        b.type = int(sc.bodies["type"][i])
    return arr


def arbiters_from_dump(arbs, prev_arbs, prev_hi, hi, sc):
    """Reference arbiter rows after a step -> restatement arbiters with the accumulators they STARTED the
    step with (cpArbiterUpdate copies jnAcc/jtAcc from the previous step's contact with the same hash)."""
    prev = {}
    for r, h in zip(prev_arbs, prev_hi):
        key = (int(min(r[0], r[1])), int(max(r[0], r[1])))
        prev[key] = [(int(r[12 + 12 * k + 11]) | (int(h[k]) << 32), r[12 + 12 * k + 7], r[12 + 12 * k + 8]) for k in range(int(r[2]))]
    out = (Arbiter * max(len(arbs), 1))()
    for i, (r, h) in enumerate(zip(arbs, hi)):
        a = out[i]
        a.body_a = int(sc.shapes["body"][int(r[0])]); a.body_b = int(sc.shapes["body"][int(r[1])])
        a.count = int(r[2]); a.first_collision = int(r[3] == 0)
        a.n = Vec(r[4], r[5]); a.e = r[6]; a.u = r[7]; a.surface_vr = Vec(r[8], r[9])
        old = prev.get((int(min(r[0], r[1])), int(max(r[0], r[1]))), [])
        for k in range(a.count):
            q = r[12 + 12 * k: 24 + 12 * k]
            c = a.contacts[k]
            c.r1 = Vec(q[0], q[1]); c.r2 = Vec(q[2], q[3])
            hk = int(q[11]) | (int(h[k]) << 32)
            c.jnAcc = 0.0; c.jtAcc = 0.0
            for (oh, ojn, ojt) in old:
                if oh == hk:
                    c.jnAcc = ojn; c.jtAcc = ojt
    return out


@pytest.mark.parametrize("name,at", [("SimpleTerrainCircles_100", 120), ("SimpleTerrainHexagons_100", 150), ("SimpleTerrainBoxes_100", 200)])
def test_whole_step_replay_is_bit_identical(ref, cpo, name, at):
    """K1 -> K8 -> K9 -> K11 of one reference step replayed by the restatement in the reference's arbiter order."""
    sc = golden_scene(name)
    rs = ref.load(sc.blob)
    dt = sc.dt
    rs.step(dt, at)
    prev_arbs, prev_hi = rs.priv_arbiters()
    before = rs.priv_bodies()
    rs.step(dt)
    arbs, hi = rs.priv_arbiters()
    after = rs.priv_bodies()
    n = len(before)
    bodies = bodies_from_priv(before, sc)
    T = (C.c_double * 6)()
    for i in range(1, n):
        if bodies[i].type != 2:
            cpo.cpo_body_update_position(C.byref(bodies[i]), dt, T)
            assert [bodies[i].p.x, bodies[i].p.y, bodies[i].a] == [after[i][0], after[i][1], after[i][4]]
            assert list(T) == list(after[i][12:18])
    A = arbiters_from_dump(arbs, prev_arbs, prev_hi, hi, sc)
    h = sc.header
    slop = float(h["collision_slop"]); bias_coef = 1.0 - float(h["collision_bias"]) ** dt
    for i in range(len(arbs)):
        cpo.cpo_arbiter_prestep(C.byref(A[i]), bodies, dt, slop, bias_coef)
        for k in range(A[i].count):
            q = arbs[i][12 + 12 * k: 24 + 12 * k]
            c = A[i].contacts[k]
            assert [c.nMass, c.tMass, c.bounce, c.bias] == [q[4], q[5], q[6], q[10]], (i, k)
    damping = float(h["damping"]) ** dt
    g = Vec(float(h["gravity"][0]), float(h["gravity"][1]))
    for i in range(1, n):
        if bodies[i].type == 0:
            cpo.cpo_body_update_velocity(C.byref(bodies[i]), g, damping, dt)
    cpo.cpo_solve(len(arbs), A, 0, None, bodies, int(h["iterations"]), dt, 1.0)
    for i in range(1, n):
        b = bodies[i]
        assert [b.v.x, b.v.y, b.w, b.v_bias.x, b.v_bias.y, b.w_bias] == [after[i][2], after[i][3], after[i][5], after[i][6], after[i][7], after[i][8]], i
    for i in range(len(arbs)):
        for k in range(A[i].count):
            q = arbs[i][12 + 12 * k: 24 + 12 * k]
            c = A[i].contacts[k]
            assert [c.jnAcc, c.jtAcc, c.jBias] == [q[7], q[8], q[9]]


def joint_scene():
    """Six free bodies tied by one joint of each class named by the north star (no shapes => no contacts)."""
    h = np.zeros((), dtype=SCENE_HEADER)
    h["iterations"] = 10; h["collision_persistence"] = 3; h["gravity"] = (0.0, -100.0); h["damping"] = 0.9
    h["sleep_time_threshold"] = np.inf; h["collision_slop"] = 0.1; h["collision_bias"] = COLLISION_BIAS_DEFAULT; h["timestep"] = 1.0 / 60.0
    b = np.zeros(7, dtype=SCENE_BODY)
    b["type"][0] = 2; b["is_space_static"][0] = 1; b["m"][0] = np.inf; b["i"][0] = np.inf
    for i in range(1, 7):
        b["m"][i] = 1.0 + 0.5 * i; b["i"][i] = 20.0 + 3.0 * i
        b["p"][i] = (12.0 * i, 5.0 * (i % 3)); b["v"][i] = (3.0 - i, 2.0 * i); b["w"][i] = 0.3 * i - 1.0; b["a"][i] = 0.1 * i
    j = np.zeros(6, dtype=SCENE_JOINT)
    j["max_force"] = np.inf; j["max_bias"] = np.inf; j["error_bias"] = ERROR_BIAS_DEFAULT; j["collide_bodies"] = 1
    j["type"] = [0, 1, 2, 4, 8, 0]
    j["a"] = [1, 2, 3, 4, 5, 0]; j["b"] = [2, 3, 4, 5, 6, 1]
    j["anchor_a"] = [(1, 0), (0, 1), (2, 2), (-1, 0), (0, 0), (0, 50)]
    j["anchor_b"] = [(-1, 0), (0, -1), (-9.5, 3), (1, 1), (0, 0), (0, 1)]
    j["prm"][0, 0] = 11.0
    j["prm"][1, 0] = 2.0; j["prm"][1, 1] = 9.0
    j["prm"][3, 0] = 8.0; j["prm"][3, 1] = 40.0; j["prm"][3, 2] = 0.7
    j["prm"][4, 0] = 0.2; j["prm"][4, 1] = 2.0
    j["prm"][5, 0] = 30.0
    j["max_force"][1] = 5000.0
    return Scene.build(h, b, np.zeros(0, dtype=SCENE_SHAPE), np.zeros((0, 2)), j)


def test_joint_step_replay_is_bit_identical(ref, cpo):
    """pin, slide, pivot, damped spring and gear: preStep + cached impulse + iterations of one reference step."""
    sc = joint_scene()
    rs = ref.load(sc.blob)
    dt = sc.dt
    rs.step(dt, 7)
    before = rs.priv_bodies()
    jprev = rs.priv_joints()
    rs.step(dt)
    after = rs.priv_bodies()
    jafter = rs.priv_joints()
    order = rs.constraint_order()
    n = len(before)
    bodies = bodies_from_priv(before, sc)
    T = (C.c_double * (6 * n))()
    T[0:6] = [1, 0, 0, 1, 0, 0]
    Ti = (C.c_double * 6)()
    for i in range(1, n):
        cpo.cpo_body_update_position(C.byref(bodies[i]), dt, Ti)
        T[6 * i: 6 * i + 6] = list(Ti)
    J = (Joint * len(order))()
    for q, idx in enumerate(order):
        s = sc.joints[int(idx)]
        jj = J[q]
        jj.type = int(s["type"]); jj.a = int(s["a"]); jj.b = int(s["b"])
        jj.maxForce = float(s["max_force"]); jj.errorBias = float(s["error_bias"]); jj.maxBias = float(s["max_bias"])
        jj.anchorA = Vec(*s["anchor_a"]); jj.anchorB = Vec(*s["anchor_b"]); jj.prm = (C.c_double * 4)(*s["prm"])
        jj.jnAcc = jprev[int(idx)][9]; jj.jAcc2 = Vec(jprev[int(idx)][9], jprev[int(idx)][10])
        cpo.cpo_joint_prestep(C.byref(jj), bodies, T, dt)
    h = sc.header
    g = Vec(float(h["gravity"][0]), float(h["gravity"][1]))
    damping = float(h["damping"]) ** dt
    for i in range(1, n):
        cpo.cpo_body_update_velocity(C.byref(bodies[i]), g, damping, dt)
    cpo.cpo_solve(0, None, len(order), J, bodies, int(h["iterations"]), dt, 1.0)
    for i in range(1, n):
        b = bodies[i]
        assert [b.v.x, b.v.y, b.w] == [after[i][2], after[i][3], after[i][5]], i
    for q, idx in enumerate(order):
        t = int(sc.joints[int(idx)]["type"])
        got = (J[q].jAcc2.x, J[q].jAcc2.y) if t == 2 else (J[q].jnAcc,)
        want = (jafter[int(idx)][9], jafter[int(idx)][10]) if t == 2 else (jafter[int(idx)][9],)
        assert got == want, (q, idx, t)


VEC2_JOINTS = (2, 3)      # pivot, groove: jAcc is a vector


def replay_joint_step(rs, cpo, sc, dt):
    """One reference step of a contact-free scene replayed through the restated preStep / cached impulse /
    iterations, in the reference's constraint order.  Returns (bodies, joints, order, after, jafter)."""
    before = rs.priv_bodies()
    jprev = rs.priv_joints()
    rs.step(dt)
    after = rs.priv_bodies()
    jafter = rs.priv_joints()
    order = rs.constraint_order()
    n = len(before)
    bodies = bodies_from_priv(before, sc)
    T = (C.c_double * (6 * n))()
    T[0:6] = [1, 0, 0, 1, 0, 0]
    Ti = (C.c_double * 6)()
    for i in range(1, n):
        cpo.cpo_body_update_position(C.byref(bodies[i]), dt, Ti)
        T[6 * i: 6 * i + 6] = list(Ti)
    J = (Joint * len(order))()
    for q, idx in enumerate(order):
        s = sc.joints[int(idx)]
        jj = J[q]
        jj.type = int(s["type"]); jj.a = int(s["a"]); jj.b = int(s["b"])
        jj.maxForce = float(s["max_force"]); jj.errorBias = float(s["error_bias"]); jj.maxBias = float(s["max_bias"])
        jj.anchorA = Vec(*s["anchor_a"]); jj.anchorB = Vec(*s["anchor_b"]); jj.prm = (C.c_double * 4)(*s["prm"])
        jj.jnAcc = jprev[int(idx)][9]; jj.jAcc2 = Vec(jprev[int(idx)][9], jprev[int(idx)][10])
        if jj.type == 7:
            jj.prm[0] = jprev[int(idx)][8]          # the ratchet's angle is solver state (cpRatchetJoint.c:40)
        cpo.cpo_joint_prestep(C.byref(jj), bodies, T, dt)
    h = sc.header
    g = Vec(float(h["gravity"][0]), float(h["gravity"][1]))
    damping = float(h["damping"]) ** dt
    for i in range(1, n):
        cpo.cpo_body_update_velocity(C.byref(bodies[i]), g, damping, dt)
    return bodies, J, order, after, jafter


def test_all_ten_joint_classes_replay_is_bit_identical(ref, cpo):
    """groove, damped rotary spring, rotary limit, ratchet and simple motor join the five above: every joint class
    of the reference, one step replayed from several points of a run (limits and ratchets engaged and free), through
    cpo_solve AND through cpo_solve_sequence (the merged-sequence loop the coloured-order replay uses)."""
    sc = all_joints_scene(shapes=False)
    rs = ref.load(sc.blob)
    dt = sc.dt
    engaged = {6: 0, 7: 0}
    for gap in (1, 6, 33, 36, 50, 94, 12):      # limit engaged from step ~74, ratchet from ~216
        rs.step(dt, gap)
        for seq in (False, True):
            if not seq:
                # replay the SAME step twice (plain loop, then the sequence loop): snapshot by stepping a twin
                twin = ref.load(sc.blob)
                twin.step(dt, int(rs.counts()["stamp"]))
                assert np.array_equal(np.nan_to_num(twin.priv_bodies()), np.nan_to_num(rs.priv_bodies()))
                target = twin
            else:
                target = rs
            bodies, J, order, after, jafter = replay_joint_step(target, cpo, sc, dt)
            n = len(after)
            if seq:
                items = (C.c_int64 * len(order))(*[-(q + 1) for q in range(len(order))])
                cpo.cpo_solve_sequence(len(order), items, None, J, bodies, int(sc.header["iterations"]), dt, 1.0)
            else:
                cpo.cpo_solve(0, None, len(order), J, bodies, int(sc.header["iterations"]), dt, 1.0)
            for i in range(1, n):
                b = bodies[i]
                assert [b.v.x, b.v.y, b.w] == [after[i][2], after[i][3], after[i][5]], (gap, seq, i)
            for q, idx in enumerate(order):
                t = int(sc.joints[int(idx)]["type"])
                got = (J[q].jAcc2.x, J[q].jAcc2.y) if t in VEC2_JOINTS else (J[q].jnAcc,)
                want = (jafter[int(idx)][9], jafter[int(idx)][10]) if t in VEC2_JOINTS else (jafter[int(idx)][9],)
                assert got == want, (gap, seq, q, idx, t)
                if t in engaged and J[q].bias != 0.0:
                    engaged[t] += 1
    assert engaged[6] > 0 and engaged[7] > 0, engaged     # both limit-type joints were exercised at their limit


def test_pair_set_restatement(ref, cpo):
    sc = golden_scene("ComplexTerrainHexagons_1000")
    rs = ref.load(sc.blob)
    rs.step(sc.dt, 90)
    bb = np.ascontiguousarray(rs.shape_bbs())
    body = np.ascontiguousarray(sc.shapes["body"], dtype=np.int32)
    active = np.ascontiguousarray(sc.bodies["type"] != 2, dtype=np.int32)
    group = np.ascontiguousarray(sc.shapes["group"], dtype=np.uint64)
    cat = np.ascontiguousarray(sc.shapes["categories"], dtype=np.uint32)
    mask = np.ascontiguousarray(sc.shapes["mask"], dtype=np.uint32)
    out = np.zeros(100000, dtype=np.uint64)
    n = cpo.cpo_pairs(len(body), bb.ctypes.data, body.ctypes.data, active.ctypes.data, group.ctypes.data, cat.ctypes.data, mask.ctypes.data,
                      0, None, len(out), out.ctypes.data)
    assert np.array_equal(out[:n], rs.pairs())
