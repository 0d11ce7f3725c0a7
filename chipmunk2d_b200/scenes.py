"""Synthetic scene generators for BASELINE.json configs 3-5 and access to the committed golden
scenes (configs 1, 2 and the demo halves of 5).  Everything is emitted as a scene blob
(scenes/cpb_scene.h), so the identical scene can be instantiated in the reference
(oracle/_ref, through scene_io.c) and in the B200 build.

Generators are deterministic: xorshift64 seeded with 88172645463325252.
"""
import math
import os

import numpy as np

from .engine import (Scene, SCENE_HEADER, SCENE_BODY, SCENE_SHAPE, SCENE_JOINT)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ALL_CATEGORIES = 0xFFFFFFFF
COLLISION_BIAS_DEFAULT = math.pow(1.0 - 0.1, 60.0)      # cpSpace.c:137
ERROR_BIAS_DEFAULT = math.pow(1.0 - 0.1, 60.0)           # cpConstraint.c:52


def golden_scene(name):
    """Scene flattened from the reference's own demo code (tests/golden/make_golden.py)."""
    path = os.path.join(GOLDEN_DIR, name + ".scene")
    with open(path, "rb") as f:
        return Scene(f.read())


def golden_names():
    return sorted(f[:-6] for f in os.listdir(GOLDEN_DIR) if f.endswith(".scene"))


class XorShift64:
    def __init__(self, seed=88172645463325252):
        self.s = np.uint64(seed)

    def uniform(self, n):
        """n doubles in [0, 1) from a vectorised xorshift64* stream (splitmix-seeded lanes)."""
        idx = np.arange(n, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = self.s + idx * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            self.s = self.s + np.uint64(n) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _header(iterations=10, gravity=(0.0, -100.0), sleep=np.inf, slop=0.5, dt=1.0 / 60.0):
    h = np.zeros((), dtype=SCENE_HEADER)
    h["iterations"] = iterations
    h["collision_persistence"] = 3
    h["gravity"] = gravity
    h["damping"] = 1.0
    h["idle_speed_threshold"] = 0.0
    h["sleep_time_threshold"] = sleep
    h["collision_slop"] = slop
    h["collision_bias"] = COLLISION_BIAS_DEFAULT
    h["timestep"] = dt
    return h


def _static_body():
    b = np.zeros(1, dtype=SCENE_BODY)
    b["type"] = 2
    b["is_space_static"] = 1
    b["m"] = np.inf
    b["i"] = np.inf
    return b


def _segments(points_a, points_b, e=1.0, u=1.0):
    n = len(points_a)
    s = np.zeros(n, dtype=SCENE_SHAPE)
    s["type"] = 1
    s["body"] = 0
    s["categories"] = ALL_CATEGORIES
    s["mask"] = ALL_CATEGORIES
    s["e"] = e
    s["u"] = u
    s["a"] = points_a
    s["b"] = points_b
    return s


def _container(width, height, seg_len=100.0):
    """Floor + two walls of static zero-radius segments no longer than seg_len."""
    a, b = [], []
    nx = int(math.ceil(width / seg_len))
    xs = np.linspace(0.0, width, nx + 1)
    for i in range(nx):
        a.append((xs[i], 0.0)); b.append((xs[i + 1], 0.0))
    ny = int(math.ceil(height / seg_len))
    ys = np.linspace(0.0, height, ny + 1)
    for i in range(ny):
        a.append((0.0, ys[i])); b.append((0.0, ys[i + 1]))
        a.append((width, ys[i])); b.append((width, ys[i + 1]))
    return np.array(a), np.array(b)


def moment_for_poly(m, verts):
    """cpMomentForPoly with zero offset (chipmunk.c:91-110)."""
    sum1 = sum2 = 0.0
    n = len(verts)
    for i in range(n):
        v1 = verts[i]; v2 = verts[(i + 1) % n]
        a = v2[0] * v1[1] - v2[1] * v1[0]
        b = (v1[0] * v1[0] + v1[1] * v1[1]) + (v1[0] * v2[0] + v1[1] * v2[1]) + (v2[0] * v2[0] + v2[1] * v2[1])
        sum1 += a * b
        sum2 += a
    return (m * sum1) / (6.0 * sum2)


def circle_pile(n, seed=88172645463325252, sleep=0.5, columns=None, radius=5.0, iterations=10, dense=False):
    """Config 4: n radius-5 circles (demo/Bench.c add_circle: m = r^2/25, e = 0, u = 0.9) on a
    jittered hexagonal grid inside a floor+walls container; sleeping enabled."""
    rng = XorShift64(seed)
    if columns is None:
        columns = max(8, int(math.sqrt(n) * 1.5))
    rows = int(math.ceil(n / columns))
    idx = np.arange(n)
    col = idx % columns
    row = idx // columns
    if dense:
        # hexagonal close packing with a 0.1 % overlap (well inside collisionSlop, so no push-out):
        # every circle touches its six neighbours from step one -- ~3 contacts per body, the settled
        # pile the 1M-circle benchmark is specified on
        pitch = 2.0 * radius * 0.999
        rowh = pitch * math.sqrt(3.0) / 2.0
        jitter = (rng.uniform(2 * n).reshape(n, 2) - 0.5) * 0.002
        width = columns * pitch + pitch * 0.5 + 2.0 * radius * 0.004 + 0.02
        height = rows * rowh + 4 * pitch
        x = radius + 0.01 + col * pitch + (row % 2) * (pitch * 0.5) + jitter[:, 0]
        y = radius + 0.01 + row * rowh + jitter[:, 1]
    else:
        pitch = 2.0 * radius * 1.05
        width = columns * pitch + pitch
        height = rows * pitch * 0.95 + 4 * pitch
        jitter = (rng.uniform(2 * n).reshape(n, 2) - 0.5) * 0.4
        x = pitch * 0.75 + col * pitch + (row % 2) * (pitch * 0.5) * 0.9 + jitter[:, 0]
        y = radius + 1.0 + row * pitch * 0.95 + jitter[:, 1]
    bodies = np.zeros(n + 1, dtype=SCENE_BODY)
    bodies[0] = _static_body()[0]
    m = radius * radius / 25.0
    bodies["m"][1:] = m
    bodies["i"][1:] = m * (0.5 * (radius * radius))           # cpMomentForCircle(m, 0, r, 0)
    bodies["p"][1:, 0] = x
    bodies["p"][1:, 1] = y
    sa, sb = _container(width, height)
    segs = _segments(sa, sb)
    circles = np.zeros(n, dtype=SCENE_SHAPE)
    circles["type"] = 0
    circles["body"] = np.arange(1, n + 1)
    circles["categories"] = ALL_CATEGORIES
    circles["mask"] = ALL_CATEGORIES
    circles["e"] = 0.0
    circles["u"] = 0.9
    circles["r"] = radius
    shapes = np.concatenate([segs, circles])
    h = _header(iterations=iterations, sleep=sleep)
    return Scene.build(h, bodies, shapes, np.zeros((0, 2)), np.zeros(0, dtype=SCENE_JOINT))


def mixed_drop(n, seed=88172645463325252, joints=True, columns=None):
    """Config 3: n bodies, one third each of Bench.c's add_circle(5), add_box(10), add_hexagon(5)
    (bevel 1.0, e = 0, u = 0.9) on a jittered grid of spacing 11 over a segment container;
    10 % of the bodies are chained in pairs by damped springs (rest 20, k 200, damping 5) and
    10 % by pivot joints at the midpoint."""
    rng = XorShift64(seed)
    if columns is None:
        columns = max(8, int(math.sqrt(n) * 1.5))
    spacing = 11.0
    rows = int(math.ceil(n / columns))
    width = columns * spacing + spacing
    height = rows * spacing + 6 * spacing
    idx = np.arange(n)
    col = idx % columns
    row = idx // columns
    jitter = (rng.uniform(2 * n).reshape(n, 2) - 0.5) * 0.4
    x = spacing + col * spacing + jitter[:, 0]
    y = 7.0 + row * spacing + jitter[:, 1]
    kind = idx % 3                                             # 0 circle, 1 box, 2 hexagon
    bevel = 1.0
    bodies = np.zeros(n + 1, dtype=SCENE_BODY)
    bodies[0] = _static_body()[0]
    bodies["p"][1:, 0] = x
    bodies["p"][1:, 1] = y
    # add_circle(5): m = 1, I = 12.5 | add_box(10): m = 1, I = m(w^2+h^2)/12 | add_hexagon(5): m = 25, I = cpMomentForPoly
    hexagon = np.array([[math.cos(-math.pi * 2.0 * i / 6.0) * (5.0 - bevel), math.sin(-math.pi * 2.0 * i / 6.0) * (5.0 - bevel)] for i in range(6)])
    m_kind = np.array([1.0, 1.0, 25.0])
    i_kind = np.array([12.5, 1.0 * (10.0 * 10.0 + 10.0 * 10.0) / 12.0, moment_for_poly(25.0, hexagon)])
    bodies["m"][1:] = m_kind[kind]
    bodies["i"][1:] = i_kind[kind]
    sa, sb = _container(width, height)
    segs = _segments(sa, sb)
    ns = len(segs)
    shapes = np.zeros(n, dtype=SCENE_SHAPE)
    shapes["body"] = np.arange(1, n + 1)
    shapes["categories"] = ALL_CATEGORIES
    shapes["mask"] = ALL_CATEGORIES
    shapes["e"] = 0.0
    shapes["u"] = 0.9
    shapes["type"] = np.where(kind == 0, 0, 2)
    shapes["r"] = np.where(kind == 0, 5.0, bevel)
    # cpBoxShapeNew(body, size - 2 bevel, ...): verts (r,b) (r,t) (l,t) (l,b) (cpPolyShape.c:233-244)
    hw = (10.0 - 2.0 * bevel) / 2.0
    box = np.array([[hw, -hw], [hw, hw], [-hw, hw], [-hw, -hw]])
    # cpPolyShapeNew runs the hexagon through cpConvexHull; store it in the hull's output order so that
    # cpPolyShapeNewRaw reproduces it: QuickHull starts from the extreme points, which for this regular
    # hexagon yields the same clockwise cycle starting at the left-most vertex.
    hexh = hull_order(hexagon)
    n_box = int(np.count_nonzero(kind == 1)); n_hex = int(np.count_nonzero(kind == 2))
    verts = np.concatenate([np.tile(box, (n_box, 1)), np.tile(hexh, (n_hex, 1))]) if (n_box + n_hex) else np.zeros((0, 2))
    vo = np.zeros(n, dtype=np.int64)
    vo[kind == 1] = 4 * np.arange(n_box)
    vo[kind == 2] = 4 * n_box + 6 * np.arange(n_hex)
    shapes["n_verts"] = np.where(kind == 1, 4, np.where(kind == 2, 6, 0))
    shapes["vert_offset"] = vo
    all_shapes = np.concatenate([segs, shapes])
    jl = []
    if joints and n >= 20:
        # pairs (i, i + 1) on the same grid row: first 10 % springs, next 10 % pivots
        cand = idx[(col % 2 == 0) & (col + 1 < columns) & (idx + 1 < n)]
        k = max(1, n // 20)
        springs = cand[:k]
        pivots = cand[k:2 * k]
        js = np.zeros(len(springs) + len(pivots), dtype=SCENE_JOINT)
        js["max_force"] = np.inf
        js["max_bias"] = np.inf
        js["error_bias"] = ERROR_BIAS_DEFAULT
        js["collide_bodies"] = 1
        q = len(springs)
        js["type"][:q] = 4
        js["a"][:q] = springs + 1
        js["b"][:q] = springs + 2
        js["prm"][:q, 0] = 20.0; js["prm"][:q, 1] = 200.0; js["prm"][:q, 2] = 5.0
        js["type"][q:] = 2
        js["a"][q:] = pivots + 1
        js["b"][q:] = pivots + 2
        # cpPivotJointNew(a, b, pivot): anchors = world pivot in each body's local frame (angle 0 => offset)
        mid = 0.5 * (bodies["p"][pivots + 1] + bodies["p"][pivots + 2])
        js["anchor_a"][q:] = mid - bodies["p"][pivots + 1]
        js["anchor_b"][q:] = mid - bodies["p"][pivots + 2]
        jl = js
    else:
        jl = np.zeros(0, dtype=SCENE_JOINT)
    h = _header(iterations=10, sleep=np.inf)
    return Scene.build(h, bodies, all_shapes, verts, jl)


def hull_order(verts):
    """Order a convex polygon the way cpConvexHull returns it for already-convex input
    (chipmunk.c:250-274: QHull seeded with the x-extremes, clockwise-positive-area winding is kept by
    cpPolyShapeInitRaw's caller).  For the convex inputs used here the cycle is preserved; only the start
    vertex moves to the minimum-x (then minimum-y) vertex."""
    v = np.asarray(verts, dtype=np.float64)
    start = min(range(len(v)), key=lambda i: (v[i][0], v[i][1]))
    # cpConvexHull emits counter-clockwise order (positive area in cpAreaForPoly's sense)
    area = 0.0
    for i in range(len(v)):
        a = v[i]; b = v[(i + 1) % len(v)]
        area += a[0] * b[1] - a[1] * b[0]
    order = [(start + k) % len(v) for k in range(len(v))]
    if area < 0:
        order = [(start - k) % len(v) for k in range(len(v))]
    return v[order]


def batched_demo_scenes(n_spaces):
    """Config 5: alternating PyramidStack / Chains spaces (golden blobs from the reference's demos)."""
    pyr = golden_scene("PyramidStack")
    chn = golden_scene("Chains")
    return [pyr if (i % 2 == 0) else chn for i in range(n_spaces)]


def all_joints_scene(damping=0.9, shapes=True):
    """A chain of 12 free bodies, each consecutive pair tied by a different joint class; two joints go to the
    static body; one body carries a circle resting on a static segment so contacts and joints share bodies."""
    h = np.zeros((), dtype=SCENE_HEADER)
    h["iterations"] = 10; h["collision_persistence"] = 3; h["gravity"] = (0.0, -100.0); h["damping"] = damping
    h["sleep_time_threshold"] = np.inf; h["collision_slop"] = 0.1; h["collision_bias"] = COLLISION_BIAS_DEFAULT
    h["timestep"] = 1.0 / 60.0
    nb = 13
    b = np.zeros(nb, dtype=SCENE_BODY)
    b["type"][0] = 2; b["is_space_static"][0] = 1; b["m"][0] = np.inf; b["i"][0] = np.inf
    for i in range(1, nb):
        b["m"][i] = 1.0 + 0.25 * i; b["i"][i] = 15.0 + 2.0 * i
        b["p"][i] = (10.0 * i, 20.0 + 4.0 * (i % 3)); b["v"][i] = (2.0 - 0.3 * i, 1.0 * i); b["w"][i] = 0.2 * i - 1.0; b["a"][i] = 0.05 * i
    types = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 2]
    j = np.zeros(len(types), dtype=SCENE_JOINT)
    j["max_force"] = np.inf; j["max_bias"] = np.inf; j["error_bias"] = ERROR_BIAS_DEFAULT; j["collide_bodies"] = 1
    j["type"] = types
    j["a"] = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0, 0]
    j["b"] = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 1, 12]
    j["anchor_a"] = [(1, 0), (0, 1), (2, 2), (-6, 1), (-1, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (10, 60), (120, 40)]
    j["anchor_b"] = [(-1, 0), (0, -1), (-8.0, 2.5), (1, 0.5), (1, 1), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0, 1), (0, 2)]
    j["prm"][0, 0] = 9.0                                           # pin dist
    j["prm"][1, :2] = (2.0, 9.0)                                   # slide min max
    j["prm"][3, :2] = (6.0, 2.0)                                   # groove: grv_b (anchor_a is grv_a)
    j["prm"][4, :3] = (8.0, 40.0, 0.7)                             # damped spring
    j["prm"][5, :3] = (0.3, 60.0, 1.5)                             # damped rotary spring
    j["prm"][6, :2] = (-0.2, 0.4)                                  # rotary limit
    j["prm"][7, :3] = (0.05, 0.0, 0.3)                             # ratchet: angle, phase, ratchet
    j["prm"][8, :2] = (0.1, 2.0)                                   # gear phase ratio
    j["prm"][9, 0] = 1.5                                           # motor rate
    j["prm"][10, 0] = 45.0
    j["max_force"][1] = 4000.0; j["max_force"][9] = 500.0; j["max_bias"][2] = 50.0
    s = np.zeros(2, dtype=SCENE_SHAPE)
    s["categories"] = 0xFFFFFFFF; s["mask"] = 0xFFFFFFFF; s["e"] = 0.2; s["u"] = 0.7
    s["type"] = [1, 0]; s["body"] = [0, 6]
    s["a"][0] = (-50, 10); s["b"][0] = (300, 10); s["r"][1] = 6.0
    if not shapes:
        s = s[:0]
    return Scene.build(h, b, s, np.zeros((0, 2)), j)


def hub_scene(n=300, hub_radius=200.0, r=2.0):
    """One big dynamic circle touched by n small dynamic circles on a ring around it: a constraint graph with one body
    of degree n (more contacts than the 63 regular colours of the device's edge colouring can separate)."""
    h = _header(gravity=(0.0, 0.0), slop=0.1)
    b = np.zeros(n + 2, dtype=SCENE_BODY)
    b[0] = _static_body()[0]
    b["m"][1] = 50.0; b["i"][1] = 50.0 * 0.5 * hub_radius * hub_radius
    b["p"][1] = (0.0, 0.0); b["v"][1] = (3.0, -2.0); b["w"][1] = 0.05
    ang = 2.0 * math.pi * np.arange(n) / n
    d = hub_radius + r - 0.3
    b["m"][2:] = 1.0; b["i"][2:] = 0.5 * r * r
    b["p"][2:, 0] = d * np.cos(ang); b["p"][2:, 1] = d * np.sin(ang)
    b["v"][2:, 0] = -5.0 * np.cos(ang); b["v"][2:, 1] = -5.0 * np.sin(ang)
    s = np.zeros(n + 1, dtype=SCENE_SHAPE)
    s["type"] = 0
    s["body"] = np.arange(1, n + 2)
    s["categories"] = ALL_CATEGORIES; s["mask"] = ALL_CATEGORIES
    s["e"] = 0.1; s["u"] = 0.6
    s["r"][0] = hub_radius; s["r"][1:] = r
    return Scene.build(h, b, s, np.zeros((0, 2)), np.zeros(0, dtype=SCENE_JOINT))
