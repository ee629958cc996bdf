#!/usr/bin/env python
"""bench.py -- body-steps/sec of the ferox world step (frStepWorld, 10 PGS iterations).

Workload (N = 1): BASELINE.json configs[1] -- 65,536 mixed circles + convex polygons (3-8 vertices)
piled in a walled box, contact-cache warm starting on; settled for --settle steps, then timed.
N > 1: batched independent worlds -- every rank steps its own replica of that world (different seed)
on its own GPU with no data-path communication ("weak" scaling); torch.distributed (NCCL) only
provides the barrier and the max-over-ranks reduction of the timings.

One JSON line on rank 0, see the contract in the task description:
  value      whole-job body-steps/s, state resident in HBM, per-step CUDA events on the stream the
             kernels run on, L2 flushed (256 MiB memset, untimed) between steps
  e2e        same metric through the public ferox.h-level calls with HOST buffers: per step the
             host applies a force to every body (H2D), calls frStepWorld and reads every body state
             back (D2H), wall clock around the loop
  roofline   the dominant kernel (k_solve_colored): SURVEY.md section 8d algorithmic bytes
             (964 B/manifold + 376 B/contact) / its mean device time, against MEASURED_PEAKS.json
  cpu_baseline  oracle/step_oracle.c (a port; 1 core) on the same settled state, a few steps

Sub-records in the same line (driver-witnessed, each a bounded run of the same build in the same process):
  config1   BASELINE configs[0]: 1,024 boxes on static ground, 1 priming + 600 timed steps
  config3   BASELINE configs[2]: 1,048,576 uniform circles, 512 x 2,048 column, right wall removed at t = 0, on a
            16,384-unit floor -- the 600 steps of the collapse, timed; Hz, overflow, per-kernel roofline
  config3_shallow  the same 1 M circles as a bed the reference's solver can hold up (the column recipe overruns it)
  config4   BASELINE configs[3]: 4,096 independent worlds x 256 bodies.  N = 1: one batch on the GPU + the CPU
            baseline (unmodified reference, one world per process on every host core).  N > 1: the worlds
            sharded round-robin over the ranks (strong scaling) AND 4,096 worlds on every rank (weak scaling)
  config5   (N > 1) BASELINE configs[4] scaled to the box: ONE pile of N x 1,048,576 bodies cut into N strips, one per
            rank, ghost bodies and per-stage velocity changes exchanged with the library's own NCCL communicator
`--skip-sub` leaves them out (profiling runs).

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libferox_ref_70000.so, compiled from
/root/reference by oracle/Makefile) on the host cores: one world per process on every core.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "body-steps/sec (frStepWorld, 10 PGS iterations)"
UNIT = "body-steps/s"
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libferox_ref_70000.so")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies", type=int, default=65536)
    ap.add_argument("--settle", type=int, default=600)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-sub", action="store_true", help="no config1/3/4/5 sub-records (profiling runs)")
    ap.add_argument("--sub-steps", type=int, default=200, help="timed steps of the config4 / config5 sub-records")
    ap.add_argument("--sub-timeout", type=int, default=300, help="seconds after which the sub-records are abandoned and the line printed without them")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"],
                    help="config2 (default, the headline): one 65,536-body mixed pile per GPU; config3: one 1M-circle column "
                         "collapse per GPU; config4: 4,096 independent 256-body worlds batched into one launch set and "
                         "sharded over the GPUs; config5: ONE pile of N x 1,048,576 bodies cut into N strips, one per GPU, "
                         "NCCL ghost-body exchange between neighbours")
    ap.add_argument("--worlds", type=int, default=4096)
    ap.add_argument("--margin", type=float, default=5.0, help="config5: ghost margin either side of a cut")
    ap.add_argument("--ghost-cap", type=int, default=8192, help="config5: ghost rows per cut")
    a = ap.parse_args()
    if a.workload in ("config3", "config5") and a.bodies == 65536:
        a.bodies = 1048576
    return a


def workload_name(n):
    return "config2: %d mixed circles + convex polygons (3-8 verts) pile + 3 static walls, cell 2.0, dt 1/60" % n


def build_bulk_world(L, spec):
    """One world from a SceneSpec through the bulk-creation call (1M bodies in well under a second)."""
    from ferox_b200 import capi, scenes
    world = L.frCreateWorld(capi.Vector2(*spec.gravity), spec.cell)
    L.frxSetWorldCapacity(world, spec.n + 16)
    shapes = (C.c_void_p * len(spec.shapes))(*[scenes.make_shape(L, s) for s in spec.shapes])
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    pos, ang = np.ascontiguousarray(spec.pos, np.float32), np.ascontiguousarray(spec.angle, np.float32)
    got = L.frxCreateBodiesEx(world, spec.n, np.ascontiguousarray(spec.body_type, np.int32), np.ascontiguousarray(spec.body_shape, np.int32),
                              shapes, pos.reshape(-1), ptr(ang), None, None, None, None)
    assert got == spec.n, (got, spec.n)

    class World:
        pass
    b = World()
    b.world, b.lib = world, L
    b.release = lambda: L.frReleaseWorld(world)
    return b


def build_config4_batch(L, n_worlds, rank, world_size):
    """BASELINE config 4: the worlds this rank owns (round-robin), batched into ONE pipeline."""
    from ferox_b200 import capi, scenes, sharding
    mine = sharding.owned_worlds(n_worlds, rank, world_size)
    specs = [scenes.small_world(sharding.world_seed(0, w)) for w in mine]
    shape_specs = specs[0].shapes                      # every small world uses the same 5 shapes
    shapes = (C.c_void_p * len(shape_specs))(*[scenes.make_shape(L, s) for s in shape_specs])
    total = sum(s.n for s in specs)
    world = L.frxCreateWorldBatch(capi.Vector2(*specs[0].gravity), specs[0].cell, len(specs))
    L.frxSetWorldCapacity(world, total + 16)
    cat = lambda f, dt: np.ascontiguousarray(np.concatenate([getattr(s, f) for s in specs]).astype(dt))
    group = np.ascontiguousarray(np.concatenate([np.full(s.n, k, np.int32) for k, s in enumerate(specs)]))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    pos, ang, vel, om = cat("pos", np.float32), cat("angle", np.float32), cat("vel", np.float32), cat("omega", np.float32)
    got = L.frxCreateBodiesEx(world, total, cat("body_type", np.int32), cat("body_shape", np.int32), shapes,
                              pos.reshape(-1), ptr(ang), ptr(vel), ptr(om), ptr(group), None)
    assert got == total, (got, total)

    class Batch:
        pass
    b = Batch()
    b.world, b.lib, b.n_worlds = world, L, len(specs)
    b.release = lambda: L.frReleaseWorld(world)
    return b


# ------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe): sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self, window=None):
        """window = (t0, t1) in time.time() seconds: only the samples taken inside it count (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if window is None or window[0] <= r[0] <= window[1]]
        where = "timed region"
        if not inside:
            inside, where = rows, "warm-up + timed region (no sample fell inside the timed region itself)"
        sm, mx = [r[1] for r in inside], [r[2] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": where}


# ------------------------------------------------------------------------------------------------
# the reference arm: unmodified reference, one world per host core
# ------------------------------------------------------------------------------------------------
def _ref_worker(args):
    n_bodies, warmup, steps, seed, q = args
    from ferox_b200 import capi, scenes
    lib = capi.load_abi(REF_LIB)
    spec = scenes.mixed_pile(n_bodies, seed=seed)
    sc = scenes.Scene(lib, spec)
    t0 = time.perf_counter()
    sc.step(1)
    t_one = time.perf_counter() - t0
    q.put(("calib", t_one))
    sc.step(max(0, warmup - 1))
    t0 = time.perf_counter()
    sc.step(steps)
    dt = time.perf_counter() - t0
    q.put(("done", spec.n, steps, dt))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_LIB):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)"}))
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    # bounded sample: the reference's query is O(FR_WORLD_MAX_OBJECT_COUNT) per body (~3.6 s per step
    # at 65,536 bodies).  Calibrate one step; if K+W steps would not finish in ~150 s, every process
    # steps a slab of the same pile with fewer bodies (throughput per body only rises for smaller
    # worlds, so this errs in the reference's favour) and the sample says so.
    n = args.bodies
    budget = 150.0
    from ferox_b200 import capi, scenes
    lib = capi.load_abi(REF_LIB)
    probe = scenes.Scene(lib, scenes.mixed_pile(n, seed=1234))
    t0 = time.perf_counter(); probe.step(1); t_one = time.perf_counter() - t0
    probe.release()
    total_steps = args.steps + args.warmup
    sample_n = n
    if t_one * total_steps > budget:
        # cost model t ~ n * (a + b * MAX_OBJECT_COUNT): dominated by the fixed 70,000-byte scan per body
        sample_n = max(1024, int(n * budget / (t_one * total_steps)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ref_worker, args=((sample_n, args.warmup, args.steps, 1234 + i, q),)) for i in range(procs)]
    t_start = time.perf_counter()
    for p in ps:
        p.start()
    done = []
    while len(done) < procs:
        msg = q.get()
        if msg[0] == "done":
            done.append(msg)
    for p in ps:
        p.join()
    wall = max(d[3] for d in done)
    bodies = sum(d[1] for d in done)
    value = bodies * args.steps / wall
    sample = ("%d processes (one world per host core), each a %d-body world of the config-2 recipe stepped from its "
              "initial lattice (the reference cannot settle 600 steps in bounded time), %d timed steps; full-size step "
              "calibrated at %.2f s" % (procs, done[0][1], args.steps, t_one))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.bodies), "parallelism": "%d host processes" % procs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    sys.stdout.flush()
    _ = t_start



# ------------------------------------------------------------------------------------------------
# sub-records: the other BASELINE configs, measured by the same process after the headline workload
# ------------------------------------------------------------------------------------------------
def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(csv_name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes, one launch) of `kernel` from a committed ncu capture
    (profiles/summarize.py layout: a header row, a unit row, one row per launch)."""
    cap = os.path.join(ROOT, "profiles", csv_name)
    if not os.path.exists(cap):
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rows = [r.rstrip("\n").split(",") for r in open(cap)]
    try:
        hdr, unit = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for f in rows[2:]:
            if kernel in f[0]:
                return (float(f[ir]) * scale[unit[ir]] + float(f[iw]) * scale[unit[iw]],
                        "profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum, one launch)" % csv_name)
    except (ValueError, IndexError, KeyError):
        pass
    return None, None


def solver_roofline(ctr, solve_ms, kernel, traffic_csv=None):
    """SURVEY.md section 8d, solve-stage terms: 964 B per manifold + 376 B per contact per launch."""
    peak, peak_src = hbm_peak()
    M, Q = ctr["manifolds"], ctr["contacts"]
    nbytes = 964.0 * M + 376.0 * Q
    achieved = nbytes / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
    traffic, src = ncu_traffic(traffic_csv, kernel) if traffic_csv else (None, None)
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": src, "peak_source": peak_src, "bytes_per_launch": nbytes, "ms_per_launch": solve_ms}


def timed_world_steps(L, world, dt, steps, flush):
    """`steps` frStepWorld calls, each in its own CUDA-event bracket on the world's stream (L2 flushed outside the
    bracket when asked), with the per-stage events on: (per-step ms, stage means, counters, launches)."""
    import ferox_b200 as fb
    ms = np.zeros(steps, np.float32)
    L.frxSetProfiling(world, 1)
    l0 = L.frxGetLaunchCount()
    L.frxTimedSteps(world, dt, steps, 1 if flush else 0, ms)
    launches = L.frxGetLaunchCount() - l0
    stage = (C.c_double * 5)()
    k = L.frxGetStageTimes(world, stage)
    L.frxSetProfiling(world, 0)
    ctr = fb.step_counters(L, world)
    names = ["broadphase", "narrowphase", "cache", "solve", "integrate"]
    return ms, {n: stage[i] / max(1, k) for i, n in enumerate(names)}, ctr, int(launches)


def sub_config1(L, dt):
    """configs[0]: 128 x 8 boxes + ground, cell 1.5; 1 priming step (bodies enter) + 600 timed steps from the drop."""
    import ferox_b200 as fb
    from ferox_b200 import scenes
    spec = scenes.box_stack(128, 8)
    sc = scenes.Scene(L, spec, prime=False)
    L.frStepWorld(sc.world, dt)
    ms, stage, ctr, launches = timed_world_steps(L, sc.world, dt, 600, True)
    st = fb.body_states(L, sc.world)
    fb.check_error(L)
    out = {"workload": "config1: 1,024 boxes (128 x 8) on static ground, cell 1.5, dt 1/60, 1 priming + 600 timed steps, L2 flushed between steps",
           "bodies": spec.n, "steps": 600, "ms_per_step": float(ms.mean()), "value": spec.n * 600 / (float(ms.sum()) * 1e-3), "unit": UNIT,
           "stage_ms_per_step": stage, "counters": ctr, "gpu_launches": launches,
           "kinetic_energy_proxy": float((st["vel"].astype(np.float64) ** 2).sum()), "max_y": float(st["pos"][:, 1].max()),
           "roofline": solver_roofline(ctr, stage["solve"], "k_solve_local (block-local solver: the world fits one CTA)")}
    sc.release()
    return out


def sub_config3(L, dt, bodies=1048576):
    """configs[2] as SURVEY.md section 8d specifies it: the column COLLAPSE itself is the timed region."""
    import ferox_b200 as fb
    from ferox_b200 import scenes
    spec = scenes.column_collapse(bodies)
    sc = build_bulk_world(L, spec)
    L.frStepWorld(sc.world, dt)                      # priming step: queued bodies enter the world
    n = L.frGetBodyCountInWorld(sc.world)
    out = {"workload": "config3: %d uniform circles (r 0.5), column 512 wide x %d high (pitch 1.05, jitter 0.02, seed 1), left wall, right "
                       "wall removed at t = 0, static floor 16,384 units wide, cell 2.0, dt 1/60; the 600 steps of the collapse are the "
                       "timed region (per-step working set > 2 GB, far larger than the 126 MB L2)" % (bodies, (bodies + 511) // 512),
           "bodies": n, "steps": 600}
    ms_all, marks, errors, calm = [], {}, [], None
    for chunk in range(6):
        ms, stage, ctr, launches = timed_world_steps(L, sc.world, dt, 100, False)
        ms_all.append(ms)
        if not ctr["overflow"] and not L.frxGetLastError():
            calm = (ctr, stage, "steps %d-%d" % (100 * chunk + 1, 100 * (chunk + 1)))     # rooflines: the last chunk in which no step was refused
        marks["step_%d" % (100 * (chunk + 1))] = {"ms_per_step": float(ms.mean()), "counters": {k: ctr[k] for k in ("cellEntries", "candidates", "manifolds", "contacts", "colors", "overflow")},
                                                  "stage_ms_per_step": stage}
        if L.frxGetLastError():                      # reported, then the run goes on: the next step re-measures and recovers
            errors.append("steps %d-%d: %s" % (100 * chunk + 1, 100 * (chunk + 1), L.frxGetLastErrorString().decode()))
            L.frxClearLastError()
    ms_all = np.concatenate(ms_all)
    st = fb.body_states(L, sc.world)
    speed = np.linalg.norm(st["vel"], axis=1)
    out.update({"steps": int(len(ms_all)), "ms_per_step": float(ms_all.mean()), "ms_per_step_max": float(ms_all.max()),
                "hz": 1e3 / float(ms_all.mean()), "hz_worst_step": 1e3 / float(ms_all.max()),
                "value": n * len(ms_all) / (float(ms_all.sum()) * 1e-3), "unit": UNIT,
                "overflow": int(ctr["overflow"]), "error": errors or None,
                "finite": bool(np.isfinite(st["pos"]).all() and np.isfinite(st["vel"]).all()),
                "mean_speed": float(speed.mean()), "max_speed": float(speed.max()),
                "extent_x": [float(st["pos"][:, 0].min()), float(st["pos"][:, 0].max())],
                "extent_y": [float(st["pos"][:, 1].min()), float(st["pos"][:, 1].max())],
                "history": marks, "gpu_launches_last_100": launches,
                "note": "the reference's own arithmetic does not hold this column up: 10 sweeps per step cannot carry 2,000 rows, the column "
                        "telescopes into the floor in free fall and bursts later (oracle beside the GPU, bit for bit on a 16-column slice: "
                        "tests/test_gpu_large.py::test_config3_column_recipe_overruns_the_reference_solver_too); steady-state 1 M-body "
                        "figures: config3_shallow",
                "roofline_at": calm[2] if calm else None,
                "roofline": solver_roofline(calm[0], calm[1]["solve"], "k_solve_colored") if calm else None,
                "roofline_narrowphase": narrow_roofline(calm[0], calm[1]["narrowphase"]) if calm else None,
                "roofline_broadphase": broad_roofline(n, calm[0], calm[1]["broadphase"]) if calm else None})
    L.frxClearLastError()
    sc.release()
    return out


def sub_config3_shallow(L, dt, bodies=1048576, steps=200):
    """The same 1,048,576 uniform circles as a bed the reference's solver CAN hold up (8 storeys x 16 lattice rows),
    settled 600 steps: the steady-state figure for a 1 M-body world, beside the column recipe that overruns the
    solver (tests/test_gpu_large.py::test_config3_column_recipe_overruns_the_reference_solver_too)."""
    import ferox_b200 as fb
    from ferox_b200 import scenes
    spec, _, _ = scenes.strip_pile(bodies, 0, 1, seed=1, storeys=8, max_rows=16, circle_fraction=1.0)
    sc = build_bulk_world(L, spec)
    L.frStepWorld(sc.world, dt)
    n = L.frGetBodyCountInWorld(sc.world)
    L.frxStepWorldN(sc.world, dt, 600)
    L.frxSynchronizeWorld(sc.world)
    ms, stage, ctr, launches = timed_world_steps(L, sc.world, dt, steps, False)
    st = fb.body_states(L, sc.world)
    fb.check_error(L)
    speed = np.linalg.norm(st["vel"], axis=1)
    out = {"workload": "config3_shallow: %d uniform circles (r 0.5) on 8 storeys x 16 lattice rows, walled, cell 2.0, dt 1/60, settled 600 "
                       "steps (per-step working set > 2 GB, far larger than the 126 MB L2)" % bodies,
           "bodies": n, "steps": steps, "ms_per_step": float(ms.mean()), "ms_per_step_median": float(np.median(ms)), "ms_per_step_max": float(ms.max()),
           "hz": 1e3 / float(ms.mean()), "value": n * steps / (float(ms.sum()) * 1e-3),
           "unit": UNIT, "stage_ms_per_step": stage, "counters": ctr, "gpu_launches": launches, "overflow": int(ctr["overflow"]),
           "mean_speed": float(speed.mean()), "max_speed": float(speed.max()),
           "roofline": solver_roofline(ctr, stage["solve"], "k_solve_colored", "r02_config3_top_kernels_ncu.csv"),
           "roofline_narrowphase": narrow_roofline(ctr, stage["narrowphase"]),
           "roofline_broadphase": broad_roofline(n, ctr, stage["broadphase"])}
    sc.release()
    return out


def narrow_roofline(ctr, ms):
    peak, src = hbm_peak()
    nbytes = ctr["candidates"] * (64.0 + 8.0) + 96.0 * ctr["manifolds"]      # C (64 + G) + 96 M, circles: G = 8
    a = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"bound": "hbm", "kernel": "k_narrowphase_tiled", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
            "bytes_per_launch": nbytes, "ms_per_launch": ms, "peak_source": src}


def broad_roofline(n, ctr, ms):
    peak, src = hbm_peak()
    K, Cn = ctr["cellEntries"], ctr["candidates"]
    nbytes = n * (52.0 + 36.0 + 4.0) + 8.0 * K + 68.0 * K + 28.0 * K + 8.0 * Cn     # SURVEY 8d break-down, S = 4
    a = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"bound": "hbm", "kernel": "broadphase stage (prepass, entry emission, radix sort passes, pair emission)", "achieved": a,
            "peak": peak, "unit": "GB/s", "frac": a / peak, "bytes_per_launch": nbytes, "ms_per_launch": ms, "peak_source": src}


def build_config5_strip(L, per_strip, rank, world_size, local_rank, dist, margin, ghost_cap):
    """This rank's strip of ONE pile of world_size x per_strip bodies; neighbours linked through the library's own NCCL
    communicator (the 128-byte id travels over torch.distributed)."""
    from ferox_b200 import scenes
    spec, lo, hi = scenes.strip_pile(per_strip, rank, world_size, storeys=8)
    sc = scenes.StripScene.local(L, spec, rank, world_size, lo, hi, margin=margin, ghost_cap=ghost_cap, device=local_rank)
    sc.world = sc.worlds[0]
    if world_size > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            assert L.frxStripGetNcclUniqueId(buf) == 0, L.frxGetLastErrorString()
            uid = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        uid = uid.to(torch.device("cuda", local_rank))
        dist.broadcast(uid, 0)
        assert L.frxStripInitNccl(sc.world, rank, world_size, uid.cpu().numpy().tobytes()) == 0, L.frxGetLastErrorString()
    sc.step(1)
    return sc, spec


def measure_config4(L, dt, n_worlds, rank, world_size, steps, barrier, reduce_max):
    """`n_worlds` independent 256-body worlds sharded round-robin over `world_size` ranks; returns (total bodies of this
    rank, elapsed ms of the timed steps MAX-reduced, stage means, counters)."""
    sc = build_config4_batch(L, n_worlds, rank, world_size)
    L.frStepWorld(sc.world, dt)
    n = L.frGetBodyCountInWorld(sc.world)
    L.frxStepWorldN(sc.world, dt, 300)
    L.frxSynchronizeWorld(sc.world)
    barrier()
    ms, stage, ctr, launches = timed_world_steps(L, sc.world, dt, steps, True)
    barrier()
    elapsed = reduce_max(float(ms.sum()))
    sc.release()
    return n, elapsed, stage, ctr, launches, sc.n_worlds


def _ref_small_world_worker(args):
    k, steps, q = args
    from ferox_b200 import capi, scenes
    lib = capi.load_abi(os.path.join(ROOT, "oracle", "_ref", "libferox_ref_2048.so"))
    sc = scenes.Scene(lib, scenes.small_world(k))
    sc.step(300)
    t0 = time.perf_counter()
    sc.step(steps)
    q.put((sc.spec.n, steps, time.perf_counter() - t0))


def config4_cpu_baseline(steps=200):
    """the UNMODIFIED reference, one 256-body world per process on every host core."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libferox_ref_2048.so")
    if not os.path.exists(lib):
        return None
    cores = max(1, min(os.cpu_count() or 1, 64))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ref_small_world_worker, args=((k, steps, q),)) for k in range(cores)]
    for p in ps:
        p.start()
    done = [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = max(d[2] for d in done)
    return {"value": sum(d[0] for d in done) * steps / wall, "unit": UNIT, "cores": cores, "kind": "reference",
            "world_steps_per_s": cores * steps / wall,
            "sample": "%d processes, each the unmodified reference (FR_WORLD_MAX_OBJECT_COUNT 2048) stepping one config-4 world "
                      "(256 bodies, settled 300 steps) %d times; aggregate over the processes, slowest process's wall clock" % (cores, steps)}


# ------------------------------------------------------------------------------------------------
# the product arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world_size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import ferox_b200 as fb
    from ferox_b200 import scenes
    L = fb.lib()
    if L.frxSetDevice(local_rank) != 0:
        raise SystemExit("no CUDA device %d: the ferox world step has no CPU fallback" % local_rank)

    dt = scenes.DT
    strips = args.workload == "config5"

    def barrier():
        if dist is not None:
            dist.barrier()

    if args.workload == "config4":
        sc = build_config4_batch(L, args.worlds, rank, world_size)
        L.frStepWorld(sc.world, dt)
        n = L.frGetBodyCountInWorld(sc.world)
    elif strips:
        sc, spec = build_config5_strip(L, args.bodies, rank, world_size, local_rank, dist, args.margin, args.ghost_cap)
        n = L.frGetBodyCountInWorld(sc.world)
        assert n == spec.n, (n, spec.n)
    else:
        if args.workload == "config3":
            # uniform circles, 8 storeys x 16 lattice rows: deep beds of equal circles "boil" in the reference's own
            # arithmetic (oracle: a 96-row bed keeps mean |v| ~10 for hundreds of steps) and at 1 M bodies some body
            # eventually runs away, so the 1 M-circle workload is kept shallow
            spec, _, _ = scenes.strip_pile(args.bodies, 0, 1, seed=1 + rank, storeys=8, max_rows=16, circle_fraction=1.0)
        else:
            spec = scenes.mixed_pile(args.bodies, seed=1234 + rank)
        sc = scenes.Scene(L, spec, prime=False) if args.workload == "config2" else build_bulk_world(L, spec)
        L.frStepWorld(sc.world, dt)                  # priming step: queued bodies enter the world
        n = L.frGetBodyCountInWorld(sc.world)
        assert n == spec.n, (n, spec.n)

    def step_n(k):
        if strips:
            sc.step(k)
        else:
            L.frxStepWorldN(sc.world, dt, k)

    step_n(args.settle)                              # settle the pile (untimed)
    L.frxSynchronizeWorld(sc.world)
    fb.check_error(L)

    warm = max(3, args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()                                  # nvidia-smi needs a few hundred ms before its first sample
    step_n(warm)
    L.frxSynchronizeWorld(sc.world)

    # ---- timed region A: device-resident state, per-step CUDA events, cold L2 ----
    ms = np.zeros(args.steps, np.float32)
    L.frxSetProfiling(sc.world, 1)
    launches0 = L.frxGetLaunchCount()
    barrier()
    t_region0 = time.time()
    ncu_range = os.environ.get("FEROX_NCU_RANGE") == "1"     # `ncu --profile-from-start off` captures only this span
    if ncu_range:
        L.frxProfilerRange(1)
    if strips:
        # per-step brackets would serialise the ranks; one bracket around K steps on this strip's stream.
        # The per-step working set (> 2 GB for 1M bodies) is far larger than the 126 MB L2.
        L.frxRecordEvent(sc.world, 0)
        sc.step(args.steps)
        L.frxRecordEvent(sc.world, 1)
        ms[:] = float(L.frxEventElapsedMs(sc.world, 0, 1)) / args.steps
    else:
        L.frxTimedSteps(sc.world, dt, args.steps, 1, ms)
    if ncu_range:
        L.frxProfilerRange(0)
    L.frxSynchronizeWorld(sc.world)
    clocks = sampler.stop((t_region0 - 0.02, time.time() + 0.02))
    barrier()
    launches = L.frxGetLaunchCount() - launches0
    stage_ms = (C.c_double * 5)()
    prof_steps = L.frxGetStageTimes(sc.world, stage_ms)
    L.frxSetProfiling(sc.world, 0)
    ctr = fb.step_counters(L, sc.world)
    fb.check_error(L)
    elapsed_ms = float(ms.sum())

    # back-to-back (warm L2, no per-step bracket) for context
    L.frxRecordEvent(sc.world, 0)
    step_n(args.steps)
    L.frxRecordEvent(sc.world, 1)
    warm_ms = float(L.frxEventElapsedMs(sc.world, 0, 1))

    # ---- timed region B: end to end through host buffers ----
    force = np.zeros((n, 2), np.float32)
    force[:, 0] = 0.01
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    st = fb.body_states(L, sc.world)
    e2e_steps = max(10, min(args.steps, 100))
    f32p = C.POINTER(C.c_float)
    vp, va, vv, vb, fmap = f32p(), f32p(), f32p(), f32p(), f32p()
    L.frxViewBodyStates(sc.world, C.byref(vp), C.byref(va), C.byref(vv), C.byref(vb))
    inv_mass = np.ctypeslib.as_array(vv, shape=(n, 4))[:, 3].copy()
    force4 = np.zeros((n, 4), np.float32)                                  # (fx, fy, torque, 0) per body
    force4[:, :2] = force * (inv_mass > 0)[:, None]                        # frApplyForceToBody ignores immovable bodies
    fview = [None, None]

    def e2e_loop(copy_out):
        """per step: forces from a host array (H2D), one step, every body state back in host memory (D2H)."""
        barrier()
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(e2e_steps):
            if copy_out:
                L.frxApplyForces(sc.world, ptr(force), None, n)        # caller array -> pinned mirror, H2D in the step
            else:
                L.frxMapBodyForces(sc.world, C.byref(fmap))            # the caller accumulates in the pinned mirror itself
                addr = C.cast(fmap, C.c_void_p).value
                if addr != fview[0]:
                    fview[0], fview[1] = addr, np.ctypeslib.as_array(fmap, shape=(n, 4))
                np.add(fview[1], force4, out=fview[1])
            if strips:
                sc.step(1)
            else:
                L.frStepWorld(sc.world, dt)
            if copy_out:
                L.frxGetBodyStates(sc.world, ptr(st["pos"]), ptr(st["angle"]), ptr(st["sincos"]), ptr(st["vel"]),
                                   ptr(st["omega"]), ptr(st["aabb"]))   # D2H + re-layout into the caller's arrays
                acc += float(st["pos"][0, 1])
            else:
                # D2H of the observation (positions + rotation, velocities) into the pinned mirror; angle / AABB would
                # follow on demand
                L.frxViewBodyStates(sc.world, C.byref(vp), None, C.byref(vv), None)
                acc += float(vp[1]) + float(vv[4 * (n - 1)])
        L.frxSynchronizeWorld(sc.world)
        secs = time.perf_counter() - t0
        barrier()
        return secs, acc

    e2e_s, checksum = e2e_loop(False)
    e2e_copy_s, _ = e2e_loop(True)
    fb.check_error(L)

    # ---- the loop an EXISTING ferox program runs, through ferox.h alone and compiled C (tests/c/e2e_loop.c): one
    # frApplyForceToBody per body, frStepWorld, one frGetBodyPosition per body ----
    ferox_h = None
    loop_path = os.path.join(ROOT, "ferox_b200", "_build", "libfx_e2e_loop.so")
    if not strips and args.workload == "config2" and rank == 0 and os.path.exists(loop_path):
        loop = C.CDLL(loop_path)
        loop.fx_e2e_ferox_h_loop.restype = C.c_double
        loop.fx_e2e_ferox_h_loop.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_double)]
        chk = C.c_double()
        hsteps = max(5, min(50, args.steps))
        loop.fx_e2e_ferox_h_loop(sc.world, 2, dt, 0.01, 0.0, C.byref(chk))
        secs = loop.fx_e2e_ferox_h_loop(sc.world, hsteps, dt, 0.01, 0.0, C.byref(chk))
        ferox_h = {"value": n * hsteps / secs, "unit": UNIT, "ms_per_step": 1e3 * secs / hsteps, "steps": hsteps, "checksum": chk.value,
                   "api": "ferox.h only, compiled C (tests/c/e2e_loop.c): per step n x frApplyForceToBody(b, frGetBodyPosition(b), f), "
                          "frStepWorld, n x frGetBodyPosition; 16 B/body up, 16 B/body (positions) + 16 B/body (velocities on first getter) down"}
        fb.check_error(L)

    # ---- the same steps with a (no-op, compiled C) collision handler installed: every live manifold goes to the
    # host before and after the solve and comes back (SURVEY.md section 8d asks for this second number) ----
    handler_ms = None
    noop_path = os.path.join(ROOT, "ferox_b200", "_build", "libfx_noop_handler.so")
    if not strips and rank == 0 and os.path.exists(noop_path):
        from ferox_b200 import capi
        noop = C.CDLL(noop_path)
        fn = C.cast(noop.fx_noop_collision_event, capi.CollisionEventFunc)
        L.frSetWorldCollisionHandler(sc.world, capi.CollisionHandler(fn, fn))
        hsteps = max(3, min(20, args.steps))
        L.frStepWorld(sc.world, dt)
        L.frxSynchronizeWorld(sc.world)
        t0 = time.perf_counter()
        for _ in range(hsteps):
            L.frStepWorld(sc.world, dt)
        L.frxSynchronizeWorld(sc.world)
        handler_ms = (time.perf_counter() - t0) * 1e3 / hsteps
        noop.fx_noop_collision_event_calls.restype = C.c_ulonglong
        handler_calls = int(noop.fx_noop_collision_event_calls())
        null = capi.CollisionEventFunc()
        L.frSetWorldCollisionHandler(sc.world, capi.CollisionHandler(null, null))
        fb.check_error(L)

    # ---- max over ranks ----
    if dist is not None:
        import torch
        from ferox_b200 import sharding
        elapsed_ms, total_bodies, (warm_ms, e2e_s, solve_ms_total, e2e_copy_s) = sharding.reduce_report(
            dist, torch.device("cuda", local_rank), elapsed_ms, n, extra_max=(warm_ms, e2e_s, stage_ms[3], e2e_copy_s))
    else:
        solve_ms_total = float(stage_ms[3])
        total_bodies = n

    if rank == 0:
        value = total_bodies * args.steps / (elapsed_ms * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        M, Q = ctr["manifolds"], ctr["contacts"]
        solve_bytes = 964.0 * M + 376.0 * Q                      # SURVEY.md section 8d, solve-stage terms
        solve_ms = solve_ms_total / max(1, prof_steps)
        achieved = solve_bytes / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
        step_bytes = (n * (160 + 2 * 24) + 104 * ctr["cellEntries"] + ctr["candidates"] * (64 + 92) + 1060 * M + 376 * Q)
        kind = L.frxGetWorldSolverKernel(sc.world) if hasattr(L, "frxGetWorldSolverKernel") else 0
        roof_kernel = ("k_solve_colored" if kind == 0 else "k_solve_local" if kind < 0 else
                       "solve stage = k_solve_colored (colouring, ordering, records; full grid) + k_sweep_cluster (warm start + 10 sweeps "
                       "in one %d-CTA thread-block cluster, velocities in distributed shared memory)" % kind)
        traffic, traffic_src = None, None
        if args.workload == "config2" and args.bodies == 65536:
            # DRAM bytes of that kernel from the committed `ncu --set full` capture of this same command
            if kind > 0:
                traffic, traffic_src = ncu_traffic("r02_config2_top_kernels_ncu.csv", "k_sweep_cluster")
            else:
                traffic, traffic_src = ncu_traffic("r02_config2_top_kernels_ncu.csv", "k_solve_colored")
                if traffic is None:
                    traffic, traffic_src = ncu_traffic("r01_end_top_kernels_ncu.csv", "k_solve_colored")
        if strips:
            # the strip step runs the solver as 12 separately launched stages with an exchange after each; there
            # are no per-stage events, so the roofline line is the WHOLE step of this rank's strip
            solve_bytes, solve_ms, roof_kernel = float(step_bytes), elapsed_ms / args.steps, "whole strip step (all kernels + exchanges)"
            achieved = solve_bytes / (solve_ms * 1e-3) / 1e9
        names = {"config2": workload_name(args.bodies),
                 "config3": "config3: %d uniform circles (r 0.5) dropped from a loose lattice onto 8 storeys x 16 rows, walled, "
                            "cell 2.0, dt 1/60" % args.bodies,
                 "config4": "config4: %d independent worlds x 256 bodies (252 dynamic mixed + floor + 2 walls + kinematic paddle), "
                            "batched into one launch set per GPU" % args.worlds,
                 "config5": "config5: ONE pile of %d x %d mixed bodies on 8 storeys, cut into %d vertical strips (1 per GPU), "
                            "ghost margin %.1f" % (world_size, args.bodies, world_size, args.margin)}
        par = {"config2": "independent worlds, 1 per GPU", "config3": "independent worlds, 1 per GPU",
               "config4": "worlds sharded round-robin over the GPUs, no communication",
               "config5": "strip domain decomposition, per-stage NCCL send/recv of ghost rows / velocity changes between neighbours"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": warm,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "config4" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": names[args.workload],
                       "bodies_per_gpu": n, "settle_steps": args.settle,
                       "solver": "graph-coloured PGS, 10 iterations",
                       "parallelism": par[args.workload],
                       "l2": "per-step working set (> 2 GB) exceeds the 126 MB L2; one event bracket around the K steps" if strips else
                             "flushed between timed steps (256 MiB memset on the step's stream, outside the event bracket)",
                       "counters": ctr},
            "value_warm_l2": total_bodies * args.steps / (warm_ms * 1e-3),
            "stage_ms_per_step": {k: stage_ms[i] / max(1, prof_steps) for i, k in
                                  enumerate(["broadphase", "narrowphase", "cache", "solve", "integrate"])},
            "step_algorithmic_GBps": step_bytes / (elapsed_ms / args.steps * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": roof_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "bytes_per_launch": solve_bytes, "ms_per_launch": solve_ms},
            "e2e": {"value": total_bodies * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n * 16),
                    "d2h_bytes_per_step": int(n * 32), "d2h_bytes_per_step_copy_out": int(n * 52), "steps": e2e_steps,
                    "checksum": checksum,
                    "api": "frxMapBodyForces + frStepWorld (frxStripStep on a strip) + frxViewBodyStates (forces written into, states read from the library's pinned host mirror)",
                    "value_copy_out": total_bodies * e2e_steps / e2e_copy_s,
                    "api_copy_out": "frxApplyForces + step + frxGetBodyStates (caller-owned pageable arrays in, 6 caller-owned arrays out)",
                    "ferox_h_loop": ferox_h},
            "handler_bridge": None if handler_ms is None else {
                "ms_per_step": handler_ms, "value": n / (handler_ms * 1e-3), "unit": UNIT,
                "note": "rank 0, wall clock, no-op C pre+post handlers installed: %d callbacks over the sample" % handler_calls},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and args.workload == "config2":
            line["cpu_baseline"] = cpu_baseline(sc, n)
    barrier()
    sc.release()

    # ---- sub-records: the other BASELINE configs on the same build, same process (every rank takes part at N > 1) ----
    # The headline must survive whatever happens in there: a watchdog prints the line without the unfinished
    # sub-records and ends the process if they take longer than --sub-timeout seconds (e.g. a rank lost inside a collective).
    import threading
    printed = threading.Lock()

    def emit(extra=None):
        if not printed.acquire(blocking=False):
            return
        if rank == 0:
            if extra:
                line.update(extra)
            print(json.dumps(line))
            sys.stdout.flush()

    def give_up():
        emit({"sub_records": "not finished within %d s: left out" % args.sub_timeout})
        os._exit(0)

    if not args.skip_sub and args.workload == "config2":
        dog = threading.Timer(args.sub_timeout, give_up)
        dog.daemon = True
        dog.start()
        sub = run_sub_records(L, args, rank, world_size, local_rank, dist, barrier)
        dog.cancel()
        emit(sub)
    else:
        emit()
    last = threading.Timer(60, lambda: os._exit(0))      # the line is out: never hang in the tear-down
    last.daemon = True
    last.start()
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    last.cancel()


def run_sub_records(L, args, rank, world_size, local_rank, dist, barrier):
    import ferox_b200 as fb
    from ferox_b200 import scenes
    dt = scenes.DT
    out = {}

    def reduce_max(x):
        if dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def guarded(name, fn):
        # a sub-record must never take the headline down with it: report the failure in its place
        try:
            return fn()
        except Exception as e:                                   # noqa: BLE001
            L.frxClearLastError()
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    steps = args.sub_steps
    if world_size == 1:
        out["config1"] = guarded("config1", lambda: sub_config1(L, dt))
        out["config3"] = guarded("config3", lambda: sub_config3(L, dt))
        out["config3_shallow"] = guarded("config3_shallow", lambda: sub_config3_shallow(L, dt))

        def c4():
            n, elapsed, stage, ctr, launches, nw = measure_config4(L, dt, args.worlds, 0, 1, steps, barrier, reduce_max)
            rec = {"workload": "config4: %d independent worlds x 256 bodies (252 dynamic mixed + floor + 2 walls + kinematic paddle), ONE batch "
                               "(one launch set per step for all worlds), settled 300 steps, L2 flushed between timed steps" % nw,
                   "bodies": n, "worlds": nw, "steps": steps, "ms_per_step": elapsed / steps, "value": n * steps / (elapsed * 1e-3), "unit": UNIT,
                   "hz": 1e3 * steps / elapsed,
                   "world_steps_per_s": nw * steps / (elapsed * 1e-3), "stage_ms_per_step": stage, "counters": ctr, "gpu_launches": launches,
                   "roofline": solver_roofline(ctr, stage["solve"], "k_solve_local (one CTA per sub-world, velocities in shared memory)")}
            if not args.no_cpu_baseline:
                rec["cpu_baseline"] = config4_cpu_baseline()
            return rec
        out["config4"] = guarded("config4", c4)
        return out

    # ---- N > 1: the two splits north_star names ----
    def c4_multi():
        n, elapsed, stage, ctr, launches, nw = measure_config4(L, dt, args.worlds, rank, world_size, steps, barrier, reduce_max)
        total = reduce_sum(n)
        strong = {"worlds_total": args.worlds, "worlds_per_gpu": nw, "bodies_total": int(total), "ms_per_step": elapsed / steps,
                  "value": total * steps / (elapsed * 1e-3), "unit": UNIT, "stage_ms_per_step_rank0": stage}
        n, elapsed, stage, ctr, launches, nw = measure_config4(L, dt, args.worlds, 0, 1, steps, barrier, reduce_max)
        total = reduce_sum(n)
        weak = {"worlds_total": args.worlds * world_size, "worlds_per_gpu": nw, "bodies_total": int(total), "ms_per_step": elapsed / steps,
                "value": total * steps / (elapsed * 1e-3), "unit": UNIT, "stage_ms_per_step_rank0": stage}
        return {"workload": "config4: independent 256-body worlds, one batch per GPU, no data-path communication; strong = %d worlds "
                            "sharded round-robin over %d GPUs, weak = %d worlds on every GPU; MAX over ranks of the device time of "
                            "%d steps (L2 flushed between steps)" % (args.worlds, world_size, args.worlds, steps),
                "n_gpus": world_size, "steps": steps, "strong": strong, "weak": weak}
    out["config4"] = guarded("config4", c4_multi)

    def c5_multi():
        sc, spec = build_config5_strip(L, 1048576, rank, world_size, local_rank, dist, args.margin, args.ghost_cap)
        transport = {0: "in-process", 1: "NCCL send/recv", 2: "peer-direct (stores into the neighbour's IPC-mapped inbox over NVLink + flags; "
                     "NCCL for set-up and migration)"}.get(L.frxStripGetTransport(sc.world), "?")
        every = 200                                  # ownership migration period (steps), inside the timed region as well
        moved, mig_s, mig_calls = 0, 0.0, 0

        def run(k):
            nonlocal moved, mig_s, mig_calls
            done = 0
            while done < k:
                chunk = min(every, k - done)
                sc.step(chunk)
                done += chunk
                if chunk == every:
                    L.frxSynchronizeWorld(sc.world)
                    t0 = time.perf_counter()
                    moved += max(0, L.frxStripMigrate(sc._arr, 1))
                    sc.step(1)                       # the step after a migration re-uploads and re-sizes: part of its cost
                    L.frxSynchronizeWorld(sc.world)
                    mig_s += time.perf_counter() - t0
                    mig_calls += 1
                    done += 1
        run(600)
        L.frxSynchronizeWorld(sc.world)
        sc.step(3)
        L.frxSynchronizeWorld(sc.world)
        L.frxClearLastError()
        L.frxResetStepCounters(sc.world)             # `overflow` below: the timed region only
        barrier()
        L.frxRecordEvent(sc.world, 0)
        run(steps)
        L.frxRecordEvent(sc.world, 1)
        elapsed = reduce_max(float(L.frxEventElapsedMs(sc.world, 0, 1)))
        barrier()
        n = L.frGetBodyCountInWorld(sc.world)
        ctr = fb.step_counters(L, sc.world)
        total = reduce_sum(n)
        over = reduce_max(float(ctr["overflow"]))
        err = L.frxGetLastErrorString().decode() if L.frxGetLastError() else None
        L.frxClearLastError()
        sc.release()
        return {"workload": "config5: ONE pile of %d x 1,048,576 mixed bodies on 8 storeys cut into %d vertical strips, one per GPU; ghost "
                            "rows (96 B) once per step and 16-byte velocity changes after each of the 11 solver stages exchanged between "
                            "neighbours (transport: see `transport`); ownership migration (frxStripMigrate) every 200 "
                            "steps, also inside the timed region; settled 600 steps; one event bracket "
                            "around %d steps, MAX over ranks (per-step working set > 2 GB per GPU)" % (world_size, world_size, steps),
                "n_gpus": world_size, "nccl_ranks": world_size, "transport": transport, "bodies_total": int(total), "steps": steps, "ms_per_step": elapsed / steps,
                "hz": 1e3 * steps / elapsed, "value": total * steps / (elapsed * 1e-3), "unit": UNIT, "counters_rank0": ctr,
                "overflow_max_over_ranks": int(over), "error_rank0": err, "migration_every_steps": every,
                "migrated_bodies_all_ranks": int(reduce_sum(moved)),
                "migration_ms_per_call_rank0": 1e3 * mig_s / max(1, mig_calls),
                "migration_note": "wall clock of frxStripMigrate + the step that follows it (re-upload, re-sizing) on rank 0, minus nothing"}
    out["config5"] = guarded("config5", c5_multi)
    return out


def cpu_baseline(sc, n):
    """oracle/step_oracle.c on the settled state the GPU just stepped: a bounded sample of steps."""
    from oracle import orc
    ow = orc.OracleWorld.from_scene(sc)
    ow.step(1)                                   # first step has a cold contact cache
    steps, t0 = 0, time.perf_counter()
    while steps < 3 or (time.perf_counter() - t0 < 12.0 and steps < 200):
        ow.step(1)
        steps += 1
    dt = time.perf_counter() - t0
    c = ow.counters()
    ow.close()
    return {"value": n * steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d steps of the same settled %d-body world in oracle/step_oracle.c (sort-based broadphase, "
                      "reference arithmetic), 1 thread; counters %s" % (steps, n, c)}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
