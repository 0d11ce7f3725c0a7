#!/usr/bin/env python
"""bench.py -- throughput of the cpSpaceStep hot path on B200 (one JSON line; see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference] [--no-sub]

Workloads (BASELINE.json configs):
  pile1m          config 4 AS BASELINE WORDS IT (DEFAULT): 1 000 000 radius-5 circles in one space, hexagonal close packing
                  (every circle touches its neighbours from the first step), sleepTimeThreshold 0.5 ("sleeping islands":
                  the islands pass runs every step), iterations 10, dt 1/60.  The column is 667 rows tall and is still
                  collapsing during the 20 + K measured steps (v_rms ~ 90), as it is in the reference; nothing can fall
                  asleep in that window.  A single space does not shard: with N GPUs every rank steps its
                  own replica ("replicas only").
  pile1m_nosleep  the same with sleepTimeThreshold = infinity (round-1 headline; no islands pass).
  pile1m_shallow  1 000 000 circles 20 rows deep, 600 settle steps: kinetic energy decays, islands fall asleep.
  batch           config 5, weak: 4096 PyramidStack / Chains spaces PER GPU.
  batch_sharded   config 5 as SURVEY 8(e) words it: 4096 spaces in total, sharded 4096/N per GPU (sharding.shard_range),
                  no data-path collective; step statistics reduced over NCCL.
  c1 / c2         configs 1 / 2 (demo/Bench.c scenes, 1000 bodies): latency-bound.
  mixed100k       config 3.

A "step" is one cpSpaceStep of the whole workload.  `value` = non-static bodies x steps / device seconds with
everything resident in HBM (CUDA events on the engine's stream, max over ranks).  `e2e` = the same metric
through the C-ABI (include/cpb200.h) with HOST buffers: every step uploads a force for every body from a
page-locked host array (H2D), steps, and downloads every body's state into a host array (D2H).
`e2e_per_body_api` = the same through the per-object Chipmunk2D C API (cpBodySetForce / cpSpaceStep /
cpBodyGetPosition on every body).  The default line also carries `sub_records`: the other configs measured in the
same process (device-timed value + C-ABI e2e each), so that one driver run holds every BASELINE config.

`--impl reference` times the UNMODIFIED reference (oracle/_ref) on the host: cpSpace on one thread and cpHastySpace on
its two solver threads, on the SAME configuration (1 M bodies for the default workload) with fewer steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PILE = 1000000
N_BATCH = int(os.environ.get("CPB200_BENCH_SPACES", "4096"))   # (override: per-GPU tuning runs of the batched workload)
CPU_SAMPLE_PILE = 100000      # cpu_baseline leg of the GPU arm: a bounded sample (about 10-20 s of host work)
CPU_SAMPLE_MIXED = 20000
SETTLE = {"pile1m": 20, "pile1m_nosleep": 20, "pile1m_shallow": 600, "batch": 300, "batch_sharded": 300, "c1": 300, "c2": 300, "mixed100k": 120}
SUB_RECORDS = ["batch_sharded", "batch", "mixed100k", "c1", "c2", "pile1m_shallow"]


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def workload_config(workload, world):
    """The configuration a workload names -- identical for the GPU arm and the reference arm."""
    single = "replicas x%d (a single space does not shard)" % world
    if workload in ("pile1m", "pile1m_nosleep", "pile1m_shallow"):
        cfg = {"workload": {"pile1m": "config4_circle_pile_1M_single_space_sleeping_islands", "pile1m_nosleep": "config4_circle_pile_1M_single_space_sleeping_off",
                            "pile1m_shallow": "config4_circle_pile_1M_single_space_20_rows_settled_and_asleep"}[workload],
               "bodies_per_gpu": N_PILE, "iterations": 10, "dt": 1.0 / 60.0,
               "sleep_time_threshold": (None if workload == "pile1m_nosleep" else 0.5), "parallelism": single,
               "l2": "inputs larger than L2 (solver rows + arbiter records >> 126 MB)"}
    elif workload == "batch":
        cfg = {"workload": "config5_batched_PyramidStack_Chains_spaces_weak_4096_per_gpu", "spaces_per_gpu": N_BATCH, "spaces_total": N_BATCH * world,
               "iterations": 30, "dt": 1.0 / 180.0, "parallelism": "spaces sharded x%d, no data-path collective" % world,
               "l2": "working set fits L2 / shared memory; latency-bound configuration"}
    elif workload == "batch_sharded":
        cfg = {"workload": "config5_batched_PyramidStack_Chains_4096_spaces_sharded", "spaces_total": N_BATCH, "spaces_per_gpu": N_BATCH // world,
               "iterations": 30, "dt": 1.0 / 180.0, "parallelism": "4096 spaces sharded 4096/%d per GPU (strong), no data-path collective" % world,
               "l2": "working set fits L2 / shared memory; latency-bound configuration"}
    elif workload == "c1":
        cfg = {"workload": "config1_simple_terrain_circles_1000", "bodies_per_gpu": 1000, "iterations": 10, "dt": 1.0 / 60.0, "parallelism": single,
               "l2": "working set fits L2; latency-bound configuration"}
    elif workload == "c2":
        cfg = {"workload": "config2_complex_terrain_hexagons_1000", "bodies_per_gpu": 1000, "iterations": 10, "dt": 1.0 / 60.0, "parallelism": single,
               "l2": "working set fits L2; latency-bound configuration"}
    elif workload == "mixed100k":
        cfg = {"workload": "config3_mixed_100k_with_springs_and_pivots", "bodies_per_gpu": 100000, "iterations": 10, "dt": 1.0 / 60.0, "parallelism": single,
               "l2": "solver rows fit L2 (default caching)"}
    else:
        raise SystemExit("unknown workload %r" % workload)
    cfg["settle_steps"] = SETTLE[workload]
    return cfg


def build_scenes(workload, rank=0, world=1, sample=None):
    """Scenes of one rank.  `sample` (bodies or spaces) builds the bounded CPU sample of the same generator."""
    from chipmunk2d_b200.scenes import circle_pile, mixed_drop, batched_demo_scenes, golden_scene
    from chipmunk2d_b200.sharding import shard_range, space_kind
    if workload in ("pile1m", "pile1m_nosleep"):
        return [circle_pile(sample or N_PILE, dense=True, sleep=(np.inf if workload == "pile1m_nosleep" else 0.5))]
    if workload == "pile1m_shallow":
        n = sample or N_PILE
        return [circle_pile(n, dense=True, sleep=0.5, columns=max(8, n // 20))]
    if workload == "batch":
        return batched_demo_scenes(sample or N_BATCH)
    if workload == "batch_sharded":
        lo, hi = shard_range(sample or N_BATCH, world, rank)
        pyr, chn = golden_scene("PyramidStack"), golden_scene("Chains")
        return [pyr if space_kind(g) == "PyramidStack" else chn for g in range(lo, hi)]
    if workload == "c1":
        return [golden_scene("SimpleTerrainCircles_1000")]
    if workload == "c2":
        return [golden_scene("ComplexTerrainHexagons_1000")]
    if workload == "mixed100k":
        return [mixed_drop(sample or 100000)]
    raise SystemExit("unknown workload %r" % workload)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples that fell inside the timed region [t0, t1].  The sampler is started before the warm-up steps
        (nvidia-smi needs ~0.2 s to deliver its first line); a timed region shorter than the 100 ms sampling
        period may hold no sample, then the nearest ones around it (the GPU is under the same load in the warm-up
        before and in the profiling steps after) are used and `window` says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        good = [(t, r) for t, r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        window = "timed region"
        rows = [r for t, r in good if t0 is None or (t0 <= t <= t1 + 0.05)]
        if not rows and good:
            mid = 0.5 * (t0 + t1)
            rows = [r for t, r in sorted(good, key=lambda tr: abs(tr[0] - mid))[:2]]
            window = "nearest samples (timed region shorter than the sampling period)"
        sm = [float(r[1]) for r in rows]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "window": window}


def algorithmic_bytes_solver(n_contacts, n_joints, iterations):
    """384 B per contact-iteration + 192 B per contact of warm start (+ ~300 B per joint pass)."""
    return 384.0 * iterations * n_contacts + 192.0 * n_contacts + 300.0 * (iterations + 1) * n_joints


# ------------------------------------------------------------------------------------------------ CPU (reference) legs

def _ref_penetration(rs):
    """Deepest contact of a reference space: -min over contacts of dot(r2 - r1 + (p_b - p_a), n) (cpArbiter.c:425)."""
    arbs, _ = rs.priv_arbiters()
    if len(arbs) == 0:
        return 0.0
    rb = rs.priv_bodies()
    sh = np.frombuffer(rs.blob.tobytes(), dtype=np.uint8)
    from chipmunk2d_b200.engine import Scene
    body = Scene(sh.tobytes()).shapes["body"]
    pa = np.nan_to_num(rb[body[arbs[:, 0].astype(int)], 0:2]); pb = np.nan_to_num(rb[body[arbs[:, 1].astype(int)], 0:2])
    n = arbs[:, 4:6]
    pen = 0.0
    for k in range(2):
        m = arbs[:, 2] > k
        q = arbs[:, 12 + 12 * k: 24 + 12 * k]
        d = np.sum((q[:, 2:4] - q[:, 0:2] + (pb - pa)) * n, axis=1)
        if np.any(m):
            pen = max(pen, float(np.max(-d[m])))
    return pen


def reference_run(*args, **kw):
    """reference_run_ on a thread with a 1 GB stack: cpSpaceProcessComponents flood-fills the contact graph recursively
    (FloodFillComponent, cpSpaceComponent.c:168-198), and a 500 000-body pile with sleeping enabled overflows the default
    8 MB stack of the unmodified reference (segmentation fault inside cpSpaceStep; 300 000 bodies still fit)."""
    box = {}

    def work():
        try:
            box["r"] = reference_run_(*args, **kw)
        except BaseException as exc:      # noqa: BLE001 -- re-raised on the caller's thread
            box["e"] = exc
    old = threading.stack_size(1 << 30)
    try:
        t = threading.Thread(target=work)
        t.start()
        t.join()
    finally:
        threading.stack_size(old)
    if "e" in box:
        raise box["e"]
    return box.get("r")


def reference_run_(workload, steps, warmup, settle, sample=None, hasty=False, threads=0):
    """The unmodified reference (oracle/_ref) stepping `workload` (or a bounded sample of its generator) on the host.
    A single space runs cpSpaceStep on one thread (or cpHastySpaceStep on its solver threads); independent spaces (the
    batched workload) run one cpSpace per host core over all cores -- ctypes releases the GIL inside the library."""
    from oracle import ref as oref
    if not oref.available():
        return None
    cores = max(1, os.cpu_count() or 1)
    scenes = build_scenes(workload, 0, 1, sample=sample)
    r = oref.Ref()
    t0 = time.perf_counter()
    spaces = [r.load(sc.blob, hasty=hasty, threads=threads) for sc in scenes]
    t_build = time.perf_counter() - t0
    dt = scenes[0].dt
    nb = sum(sc.n_dynamic() for sc in scenes)
    pool = cores if len(spaces) > 1 else 1

    def run(fn):
        if pool == 1:
            return [fn(s) for s in spaces]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=pool) as ex:
            return list(ex.map(fn, spaces))

    run(lambda s: s.step(dt, settle + warmup))
    t0 = time.perf_counter()
    per_space = run(lambda s: s.time_steps(dt, steps))
    wall = time.perf_counter() - t0
    t = wall if pool > 1 else sum(per_space)
    counts = [s.counts() for s in spaces]
    pen = _ref_penetration(spaces[0]) if (len(spaces) == 1 and nb <= 200000) else None
    for s in spaces:
        s.space = None  # leak on purpose: tearing a large reference space down is quadratic
    solver_threads = (2 if hasty else 1)     # cpHastySpace caps its solver threads at MAX_THREADS = 2 (cpHastySpace.c:387)
    how = ("cpHastySpaceStep, %d solver threads (cpHastySpaceSetThreads(%d), capped at 2 by the reference)" % (solver_threads, threads) if hasty else
           ("cpSpaceStep, 1 thread" if pool == 1 else "%d host threads, one independent cpSpace each at a time (wall clock over the pool)" % pool))
    return {"value": nb * steps / t, "unit": "body-steps/s", "cores": (pool if pool > 1 else solver_threads), "kind": "reference",
            "sample": "%s at %d bodies (%d space%s), %d settle + %d warm-up + %d timed steps, %s; gcc -O2 -ffp-contract=off no fast-math; space built in %.1f s" % (
                workload, nb, len(scenes), "" if len(scenes) == 1 else "s", settle, warmup, steps, how, t_build),
            "bodies": nb, "ms_per_step": 1000.0 * t / steps, "contacts_per_step": sum(c["contacts"] for c in counts), "arbiters_per_step": sum(c["arbiters"] for c in counts),
            "max_penetration": pen, "host_cores": cores}


def cpu_sample_size(workload):
    return {"pile1m": CPU_SAMPLE_PILE, "pile1m_nosleep": CPU_SAMPLE_PILE, "pile1m_shallow": CPU_SAMPLE_PILE, "mixed100k": CPU_SAMPLE_MIXED,
            "batch": max(64, 8 * (os.cpu_count() or 1)), "batch_sharded": max(64, 8 * (os.cpu_count() or 1))}.get(workload)


def cpu_baseline_sample(workload, with_hasty=True):
    """cpu_baseline of the GPU arm: a BOUNDED sample of the workload's generator on the box's host cores."""
    sample = cpu_sample_size(workload)
    big = workload.startswith("pile1m") or workload == "mixed100k"
    settle = (5 if big else SETTLE[workload])
    if workload == "mixed100k":
        settle = SETTLE[workload]
    steps, warm = ((10, 2) if big else (40, 10))
    base = reference_run(workload, steps, warm, settle, sample=sample)
    if base is None:
        return None
    if with_hasty and not workload.startswith("batch"):
        h = reference_run(workload, steps, warm, settle, sample=sample, hasty=True, threads=0)
        if h:
            base["hasty"] = {k: h[k] for k in ("value", "unit", "cores", "ms_per_step", "sample")}
    return base


def run_reference(args):
    rank, local_rank, world = rank_info()
    if rank != 0:
        return
    cfg = workload_config(args.workload, max(1, args.gpus))
    steps = max(1, min(args.steps, 3 if args.workload.startswith("pile1m") else 40))
    warm = max(0, min(args.warmup, 2 if args.workload.startswith("pile1m") else 10))
    settle = (5 if args.workload.startswith("pile1m") and args.workload != "pile1m_shallow" else SETTLE[args.workload])
    sample = (cpu_sample_size(args.workload) if args.workload.startswith("batch") else None)
    base = reference_run(args.workload, steps, warm, settle, sample=sample)
    if base is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"})
        return
    if not args.workload.startswith("batch"):
        big = args.workload.startswith("pile1m")
        h = reference_run(args.workload, (2 if big else steps), (1 if big else warm), (2 if big and args.workload != "pile1m_shallow" else settle), sample=sample, hasty=True, threads=0)
        if h:
            base["hasty"] = {k: h[k] for k in ("value", "unit", "cores", "ms_per_step", "sample")}
    line = {"impl": "reference", "metric": "body_steps_per_sec", "value": base["value"], "unit": "body-steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ the GPU arm

def measure(workload, args, ctx, full):
    """One workload on this rank's GPU: device-timed value, C-ABI e2e and (full) stage profile / roofline inputs."""
    import torch
    from chipmunk2d_b200.engine import World, BODY_STATE
    rank, local_rank, world, dist = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["dist"]
    scenes = build_scenes(workload, rank, world)
    cfg = workload_config(workload, world)
    dt = scenes[0].dt
    nb = sum(sc.n_dynamic() for sc in scenes)
    iterations = int(scenes[0].header["iterations"])
    steps = args.steps
    warm = max(3, args.warmup)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    w = World(len(scenes), device=local_rank)
    w.load_scenes(scenes)
    settle = SETTLE[workload]
    sampler = None
    if full:
        sampler = ClockSampler(local_rank)
        sampler.start()         # before the settle steps: nvidia-smi takes a moment to deliver its first sample
    w.step(dt, settle)          # untimed: let contacts form so the timed steps see the configured workload
    w.sync()
    if sampler is not None:
        t_wait = time.time()
        while sampler.proc and not sampler.rows and time.time() - t_wait < 2.0:
            time.sleep(0.05)    # (no extra steps here: the timed window must start at the same step in every run)
    w.step(dt, warm)
    w.sync()

    launches0 = w.launch_count()
    barrier()
    t_wall0 = time.time()
    ms = w.time_steps(dt, steps)
    barrier()
    clocks = sampler.stop(t_wall0, time.time()) if sampler is not None else None
    launches = w.launch_count() - launches0
    st = w.stats()

    # per-stage device times + solver internals (separate short pass, profiling adds a sync per step)
    w.set_profiling(True)
    acc = {}
    nprof = 5
    for _ in range(nprof):
        w.step(dt)
        for k, v in w.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / nprof
    sp = w.solver_profile()
    w.set_profiling(False)
    st2 = w.stats()

    # the only inter-GPU traffic: timings (MAX) and step statistics (SUM / MAX) -- north star
    from chipmunk2d_b200.sharding import reduce_step_stats
    sums, mx, t_ms_max = reduce_step_stats(dist, torch, "cuda", [float(nb), float(st["n_contacts"]), float(st["n_arbiters"]), float(st["n_pairs"]), st["kinetic_energy"],
                                                                    float(st["n_awake"]), float(len(scenes))], [st["max_penetration"]], ms)
    t_max = t_ms_max * 1e-3
    total_bodies, total_contacts, total_arbs, total_pairs, total_ke, total_awake, total_spaces = sums

    # ---- e2e through the C-ABI with HOST buffers every step, copies inside the timed region ----
    # `e2e`: cpb200_world_bind_io -- the host writes a force for every body into its page-locked array, cpb200_world_step,
    # cpb200_world_sync, the host finds (p, a, v, w) of every body in its page-locked array.  The engine moves the forces
    # (H2D, 24 B/body) during the collision phase and the positions (D2H, 24 B/body) during the rest of the step; only the
    # velocities (24 B/body) are copied after the solver.  `e2e_blocking`: the round-1 path, cpb200_world_set_body_forces
    # -> cpb200_world_step -> cpb200_world_get_bodies (80 B/body), every call waiting for the previous one.
    def timed_loop(body, k):
        for _ in range(3):
            body()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            body()
        sec = time.perf_counter() - t0
        t_e = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        return float(t_e.item())

    try:
        k2 = max(1, min(steps, 10))
        forces = w.pinned_array(w.n_bodies * 3, np.float64).reshape(-1, 3)   # page-locked host buffers
        forces[:] = 0.0
        states = w.pinned_array(w.n_bodies, BODY_STATE)

        def blocking():
            w.set_body_forces(0, forces); w.step(dt); w.bodies_into(states)
        sec = timed_loop(blocking, k2)
        e2e_blocking = {"value": total_bodies * k2 / sec, "unit": "body-steps/s", "steps": k2,
                        "h2d_bytes_per_step": int(w.n_bodies * 24), "d2h_bytes_per_step": int(w.n_bodies * BODY_STATE.itemsize), "ms_per_step": 1000.0 * sec / k2,
                        "path": "cpb200_world_set_body_forces -> cpb200_world_step -> cpb200_world_get_bodies (80 B/body), each call blocking"}
        sink = w.pinned_array(w.n_bodies * 6, np.float64).reshape(2, -1, 3)
        w.bind_io(forces, sink)
        checksum = [0.0]

        def bound():
            forces[0, 0] = checksum[0] * 0.0       # the host writes its input ...
            w.step(dt); w.sync()
            checksum[0] = float(sink[1, -1, 0])    # ... and reads its output, every step
        sec = timed_loop(bound, k2)
        w.bind_io(None, None)
        e2e = {"value": total_bodies * k2 / sec, "unit": "body-steps/s", "steps": k2,
               "h2d_bytes_per_step": int(w.n_bodies * 24), "d2h_bytes_per_step": int(w.n_bodies * 48), "ms_per_step": 1000.0 * sec / k2,
               "path": "C-ABI with page-locked host arrays bound to the world (cpb200_world_bind_io): forces in, cpb200_world_step, cpb200_world_sync, (p, a, v, w) of every body out -- every step; copies overlap the step's kernels"}
        e2e["blocking"] = e2e_blocking
    except Exception as exc:  # keep the device-resident number even if something on the host path is missing
        e2e = {"value": None, "unit": "body-steps/s", "error": str(exc)}

    e2e_api = None
    if full and len(scenes) == 1:
        try:
            from chipmunk2d_b200.api import load_scene_lib, SceneSpace
            k3 = max(1, min(steps, 5))
            api = SceneSpace(load_scene_lib(), scenes[0].blob)
            api.step(dt, settle)
            api.e2e_steps(dt, 2)
            barrier()
            sec, _pos = api.e2e_steps(dt, k3)
            t_a = torch.tensor([sec], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t_a, op=dist.ReduceOp.MAX)
            e2e_api = {"value": total_bodies * k3 / float(t_a.item()), "unit": "body-steps/s", "steps": k3,
                       "h2d_bytes_per_step": int(api.n_bodies * 24), "d2h_bytes_per_step": int(api.n_bodies * BODY_STATE.itemsize),
                       "ms_per_step": 1000.0 * float(t_a.item()) / k3,
                       "path": "cpBodySetForce on every body -> cpSpaceStep -> cpBodyGetPosition on every body (scene_io.c cpb_scene_e2e_steps)"}
            api.space = None
        except Exception as exc:
            e2e_api = {"value": None, "unit": "body-steps/s", "error": str(exc)}
    append_cost = None
    if full and len(scenes) == 1 and workload.startswith("pile1m"):
        # f4: what one cpSpaceAddBody + cpSpaceAddShape costs on the running space -- the appended range,
        # the first step after it (kernel by kernel: the step graph is re-captured) -- over a plain step
        try:
            from chipmunk2d_b200.engine import BODY_DESC, SHAPE_DESC
            def one_step():
                t0 = time.perf_counter(); w.step(dt); w.sync(); return time.perf_counter() - t0
            plain = min(one_step() for _ in range(5))
            costs = []
            for r in range(5):
                bd = np.zeros(1, dtype=BODY_DESC); bd["m"] = 1.0; bd["i"] = 12.5; bd["rot"][:, 0] = 1.0; bd["sleep_group"] = -1
                bd["p"] = (50.0 + 20.0 * r, 20000.0)
                sd = np.zeros(1, dtype=SHAPE_DESC); sd["body"] = w.n_bodies; sd["hashid"] = 10000000 + r; sd["r"] = 5.0
                sd["categories"] = 0xFFFFFFFF; sd["mask"] = 0xFFFFFFFF; sd["u"] = 0.9
                t0 = time.perf_counter()
                ok = w.append_bodies(bd) and w.append_shapes(sd)
                w.step(dt); w.sync()
                costs.append(time.perf_counter() - t0 - plain)
                w.step(dt, 2); w.sync()
            append_cost = {"extra_ms_per_added_body_and_shape": 1000.0 * float(np.median(costs)), "plain_step_ms": 1000.0 * plain, "appended_in_place": bool(ok),
                           "what": "cpb200_world_append_bodies(1) + cpb200_world_append_shapes(1) + the first step after them, minus a plain step (host clock, median of 5)"}
        except Exception as exc:
            append_cost = {"error": str(exc)}
    w.close()

    rec = {"config": cfg, "value": total_bodies * steps / t_max, "unit": "body-steps/s", "steps": steps, "warmup": warm,
           "ms_per_step": 1000.0 * t_max / steps,
           "contacts_solved_per_sec": total_contacts * steps / t_max,
           "contact_iterations_per_sec": total_contacts * iterations * steps / t_max,
           "per_step": {"bodies": total_bodies, "spaces": total_spaces, "pairs": total_pairs, "arbiters": total_arbs, "contacts": total_contacts, "colours": st["n_colours"],
                        "awake_bodies": total_awake, "max_penetration": float(mx[0]), "kinetic_energy": total_ke,
                        "idle_row_fraction": (st["n_row_idle"] / st["n_row_solves"] if st["n_row_solves"] else None)},
           "gpu_launches": int(launches), "e2e": e2e, "stage_us": acc, "solver_us": sp}
    if e2e_api is not None:
        rec["e2e_per_body_api"] = e2e_api
    if append_cost is not None:
        rec["append_cost"] = append_cost
    rec["_local"] = {"iterations": iterations, "nb": nb, "st2": st2, "clocks": clocks, "ms_local": ms, "n_scenes": len(scenes)}
    return rec


def roofline_of(rec, workload):
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    loc = rec["_local"]
    st2, iterations, nb = loc["st2"], loc["iterations"], loc["nb"]
    solve_us = rec["stage_us"].get("solve", 0.0)
    alg = algorithmic_bytes_solver(st2["n_contacts"], st2["n_joints"], iterations)
    achieved = alg / (solve_us * 1e-6) / 1e9 if solve_us > 0 else 0.0
    traffic = None
    dram_frac = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(workload)
    except Exception:
        pass
    if traffic and solve_us > 0:
        dram_frac = float(traffic) / (solve_us * 1e-6) / 1e9 / peak
    whole = 264.0 * nb + 108.0 * nb + 32.0 * st2["n_shapes"] + (8.0 + 170.0) * st2["n_pairs"] + 64.0 * st2["n_arbiters"] + 400.0 * st2["n_contacts"] + 384.0 * iterations * st2["n_contacts"]
    step_s = loc["ms_local"] / rec["steps"] * 1e-3
    out = {"bound": "hbm", "kernel": "k_colour_solve<..., PHASE 2> (persistent: warm start + %d Gauss-Seidel iterations + write-back; the colouring + row build is a launch of its own, stage colour_rows)" % iterations,
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
           "frac_on_dram_bytes": dram_frac,
           "peak_source": peak_kind, "algorithmic_bytes_per_launch": alg, "launch_us": solve_us,
           "whole_step": {"algorithmic_bytes": whole, "achieved_gbs": whole / step_s / 1e9, "frac": whole / step_s / 1e9 / peak}}
    if workload.startswith("batch"):
        out["note"] = "batched small spaces are solved from shared memory / L2 (space-local solver): the algorithmic bytes never reach HBM, so this fraction can exceed 1 and is not a bound for this workload"
    return out


def run_ours(args):
    rank, local_rank, world = rank_info()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the step engine has no CPU fallback)")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    os.environ["CPB200_DEVICE"] = str(local_rank)   # spaces created through the C API follow the rank's GPU
    # the host layer threads its mirror loops: share the host cores between the ranks of this node
    os.environ.setdefault("CPB200_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world))))
    ctx = {"rank": rank, "local_rank": local_rank, "world": world, "dist": dist}

    main = measure(args.workload, args, ctx, full=True)
    subs = {}
    if not args.no_sub:
        for name in SUB_RECORDS:
            if name == args.workload:
                continue
            try:
                r = measure(name, args, ctx, full=False)
                r.pop("_local", None)
                r.pop("solver_us", None)
                subs[name] = r
            except Exception as exc:
                subs[name] = {"error": str(exc)}

    if dist is not None:
        # every rank leaves the process group together, BEFORE rank 0 goes on to its CPU-only work
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    roof = roofline_of(main, args.workload)
    cpu = None
    try:
        cpu = cpu_baseline_sample(args.workload)
    except Exception as exc:
        cpu = {"value": None, "error": str(exc)}
    if "mixed100k" in subs and "error" not in subs["mixed100k"] and not args.no_sub:
        # the deepest contact of config 3 next to the reference's, same generator, same number of steps, 20 k bodies
        try:
            from chipmunk2d_b200.engine import World
            sc = build_scenes("mixed100k", sample=CPU_SAMPLE_MIXED)[0]
            wd = World(1, device=local_rank); wd.load_scene(sc); wd.step(sc.dt, SETTLE["mixed100k"] + 12); wd.sync()
            rr = reference_run("mixed100k", 10, 2, SETTLE["mixed100k"], sample=CPU_SAMPLE_MIXED)
            subs["mixed100k"]["penetration_check"] = {"bodies": CPU_SAMPLE_MIXED, "steps": SETTLE["mixed100k"] + 12, "device_max_penetration": wd.stats()["max_penetration"],
                                                      "reference_max_penetration": (rr or {}).get("max_penetration"),
                                                      "reference_body_steps_per_sec": (rr or {}).get("value")}
            wd.close()
        except Exception as exc:
            subs["mixed100k"]["penetration_check"] = {"error": str(exc)}

    loc = main.pop("_local")
    line = {
        "metric": "body_steps_per_sec", "value": main["value"], "unit": "body-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": main["warmup"], "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": ("strong" if args.workload == "batch_sharded" else "weak"), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": main["config"],
        "contacts_solved_per_sec": main["contacts_solved_per_sec"], "contact_iterations_per_sec": main["contact_iterations_per_sec"],
        "per_step": main["per_step"], "clocks": loc["clocks"], "gpu_launches": main["gpu_launches"],
        "e2e": main["e2e"], "e2e_per_body_api": main.get("e2e_per_body_api"),
        "roofline": roof, "stage_us": main["stage_us"], "solver_us": main["solver_us"],
        "append_cost": main.get("append_cost"),
        "cpu_baseline": cpu, "sub_records": subs,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else any library prints while
    the benchmark runs (e.g. NCCL's version banner) was diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=os.environ.get("CPB200_BENCH_WORKLOAD", "pile1m"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sub", action="store_true", default=(os.environ.get("CPB200_BENCH_SUB", "1") == "0"),
                    help="skip the sub-records (the other BASELINE configs measured in the same run)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
