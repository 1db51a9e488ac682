#!/usr/bin/env python
"""bench.py — DMSA iterations/s (and point-Jacobians/s) on B200 vs the CPU reference path.

A "step" is ONE DMSA iteration = the loop body DmsaOptimizer.h:69-144: base transform, two voxel-set builds,
weights, e0, the P forward-difference cost evaluations, H = J^T J, the LM step and the 9-point line search
(P + 10 cost evaluations).  Workload at N = 1: BASELINE.json configs[1] (sliding window, 10 scans x 65 536 points
+ 50 000 static points, 20 control poses).  N > 1 (default): one independent window per rank ("replicas only" for the
sliding-window pass, north star); `--workload keyframe` runs the keyframe-bundle pass with an NCCL all-reduce.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU path (oracle port) on the host cores, same JSON shape
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ["DMSA_B200_TABLE_CACHE"] = "0"  # the end-to-end leg feeds the same window every step: recompute + upload the timing tables each time
os.environ.setdefault("NCCL_DEBUG", "WARN")  # NCCL's version banner goes to stdout otherwise; stdout carries exactly one JSON line


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but native libraries (NCCL's version banner when the box exports NCCL_DEBUG)
    write to file descriptor 1 behind Python's back: keep a private duplicate of the real stdout for the result line and
    point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


RESULT_OUT = None


def emit(line):
    out = RESULT_OUT if RESULT_OUT is not None else sys.__stdout__
    out.write(json.dumps(line) + "\n")
    out.flush()

SETTINGS = dict(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30)
METRIC = "DMSA iterations/sec"
UNIT = "iterations/s"
PRODUCTION_NUM_ITER = 10  # config/slam_settings.yaml:22 (num_iter of the sliding-window optimizeSet)


def workload_string(config, win, P, world):
    """The same string in both arms (the driver compares `config` of the two lines)."""
    return (f"{config}: sliding-window DMSA iteration (DmsaOptimizer.h:69-144), {len(win['scans'])} scans x {len(win['scans'][0])} pts + "
            f"{len(win['static'])} static, {win['n_poses']} control poses, P={P}, {P + 10} cost evaluations/step; "
            + ("1 window" if world == 1 else f"{world} identical windows, one per rank (replicas, no collective)"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
            t0 = time.time()  # nvidia-smi needs a moment to attach: wait for its first sample so the timed region is covered
            while time.time() - t0 < 5.0 and os.path.getsize(self.f.name) == 0:
                time.sleep(0.05)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def best_thread_count(win):
    """OpenMP team size that runs the oracle's cost evaluation fastest on this host (cgroup quotas make nproc a poor guess)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob

    m = ob.OracleModel.from_window(win)
    m.centralize()
    m.update_global_points()
    m.build_sets(ob.settings(**SETTINGS))
    n = os.cpu_count() or 1
    cands = sorted({max(1, n >> k) for k in range(0, 6)} | {min(n, 8)})
    best, best_t = 1, float("inf")
    for c in cands:
        m.set_threads(c)
        m.time_cost_evals(1)
        t = m.time_cost_evals(2)
        if t < best_t:
            best, best_t = c, t
    return best


def oracle_cpu_baseline(win, threads, iters=2, opt="O2"):
    """The CPU path (oracle port of the reference arithmetic, oracle/dmsa_oracle.cpp) timed on the host cores.
    opt="O1", threads=1: the "reference-faithful" build (CMakeLists.txt:15 -O1; the reference's cost loops are serial,
    DmsaOptimizer.h:56-57 only gives Eigen's J^T J product 4 threads)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob

    m = ob.OracleModel.from_window(win, opt=opt)
    m.set_threads(threads)
    m.set_mode(0)
    m.centralize()
    st = ob.settings(**SETTINGS)
    p0 = m.get_params()
    t, status = m.time_iteration(st)  # warm (page faults)
    ts = []
    for _ in range(iters):
        m.set_params(p0)
        t, status = m.time_iteration(st)
        ts.append(t)
    s = m.sets()
    return dict(seconds_per_iteration=float(np.median(ts)), M=int(s["M"]), G=int(s["G"]), status=status, threads=threads)


def keyframe_cpu_baseline(sm, n_bundles, settings):
    """The CPU path (oracle port) on ONE bundle of the same submap, scaled by the number of bundles (bounded sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from dmsa_lidar_slam_b200.distributed import bundle_ranges

    f, l = bundle_ranges(sm["n_keyframes"], 15, 8)[0]
    sub = dict(n_keyframes=l - f + 1, clouds=sm["clouds"][f:l + 1], rings=sm["rings"][f:l + 1], grid_sizes=sm["grid_sizes"][f:l + 1],
               rel_orient=sm["rel_orient"][:, f:l + 1].copy(), rel_transl=sm["rel_transl"][:, f:l + 1].copy())
    m = ob.OracleModel.from_submap(sub)
    threads = os.cpu_count() or 1
    m.set_threads(threads)
    m.set_mode(0)
    t, status = m.time_iteration(ob.settings(**settings))
    return {"value": 1.0 / (t * n_bundles), "unit": "iterations/s", "cores": threads, "kind": "port",
            "sample": f"one iteration of one of the {n_bundles} bundles (15 keyframes) by oracle/dmsa_oracle.cpp, faithful arithmetic; time x {n_bundles}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dmsa_lidar_slam_b200 import synth

    win = synth.make_config(args.config)
    threads = best_thread_count(win)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob

    m = ob.OracleModel.from_window(win)
    m.set_threads(threads)
    m.set_mode(0)
    m.centralize()
    st = ob.settings(**SETTINGS)
    p0 = m.get_params()
    for _ in range(max(1, min(args.warmup, 2))):
        m.set_params(p0)
        m.time_iteration(st)
    t_total = 0.0
    steps = args.steps  # each step is one full CPU iteration of the same workload (~0.3 s at cfg2 on 16 threads)
    for _ in range(steps):
        m.set_params(p0)
        t, _ = m.time_iteration(st)
        t_total += t
    s = m.sets()
    P = m.P
    val = steps / t_total
    world = max(1, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 point arithmetic / f64 pose chain, sums, J^T J (the reference's types)", "data": "synthetic",
        "config": {"workload": workload_string(args.config, win, P, world), "N": int(m.N), "M": int(s["M"]), "G": int(s["G"]),
                   "l2": "flushed between timed steps (256 MiB write, untimed)", "settings": SETTINGS},
        "point_jacobians_per_s": val * int(s["M"]),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} full iterations of one window; oracle/dmsa_oracle.cpp -O2 -fopenmp (the reference needs Eigen/PCL/Boost: unbuildable here)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def measure_frontend(traj, win, local):
    """SURVEY 8(f) ranks 2 and 3 through the C-ABI with host buffers (every call uploads its cloud and reads its result back),
    next to the oracle port on the host cores; `identical` = the results of the two are equal."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    from dmsa_lidar_slam_b200 import PreProcessor, PreprocessConfig
    from dmsa_lidar_slam_b200.synth import POINT_NORMAL

    def gpu_ms(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        return float(np.median(ts)), r

    def cpu_ms(fn):
        t0 = time.perf_counter()
        r = fn()
        return 1e3 * (time.perf_counter() - t0), r

    out = {"what": "the producers and the neighbour of the optimizer's inputs (DmsaSlam.h:264-414, 557-634; helpers.h:67-182), host buffers in and out, "
                   "wall time per call incl. copies; cpu = oracle port, OpenMP on all host cores where the step is parallel"}
    pre = PreProcessor(ctx=traj.ctx)
    raw = win["scans"][0]
    cfg = PreprocessConfig(3000, 30.0, 0.0)
    g_ms, (g_out, g_gs) = gpu_ms(lambda: pre.preProcess(raw, cfg, 7))
    c_ms, (c_out, c_gs) = cpu_ms(lambda: ob.preprocess(raw, 3000, 30.0, 0.0, np.eye(4), 7))
    out["preprocess_scan"] = {"points_in": int(len(raw)), "points_out": int(len(g_out)), "grid_size": g_gs, "gpu_ms": g_ms, "cpu_ms": c_ms,
                              "identical": bool(g_out.tobytes() == c_out.tobytes())}
    # keyframe cloud: the window's globalPoints downsampled at minGridSize (DmsaSlam.h:506), then k = 6 normals
    traj.updateGlobalPoints()
    world = traj.globalPoints()
    import ctypes as C
    idx = np.zeros(len(world), dtype=np.int32)
    n_out = C.c_int64(0)

    def ds():
        traj.ctx._ck(traj.L.dmsa_b200_downsample_global_points(traj.h, C.c_float(0.3), 5, idx.ctypes.data_as(C.c_void_p), C.byref(n_out)))
        return idx[: n_out.value].copy()

    g_ms, g_idx = gpu_ms(ds)
    c_ms, c_idx = cpu_ms(lambda: ob.grid_downsample(world[:, :3], 0.3, 5))
    out["keyframe_downsample"] = {"points_in": int(len(world)), "points_out": int(len(g_idx)), "gpu_ms": g_ms, "cpu_ms": c_ms, "cpu_cores": 1,
                                  "identical": bool(np.array_equal(g_idx, c_idx))}
    kc = np.zeros(len(g_idx), dtype=POINT_NORMAL)
    kc["x"], kc["y"], kc["z"], kc["w"] = world[g_idx, 0], world[g_idx, 1], world[g_idx, 2], 1.0
    g_ms, (g_cl, g_nn) = gpu_ms(lambda: pre.updateNormals(kc, (0, 0, 0), 0.3, with_neighbours=True))
    sub = kc[:20000]  # the oracle's neighbour search is exhaustive (O(n^2)): bounded sample
    c_ms, (c_cl, c_nn) = cpu_ms(lambda: ob.update_normals(sub))
    g_sub = pre.updateNormals(sub, (0, 0, 0), 0.3, with_neighbours=True)
    out["update_normals_k6"] = {"points": int(len(kc)), "gpu_ms": g_ms, "points_per_s": len(kc) / (g_ms * 1e-3),
                                "cpu_ms_20000_points_exhaustive": c_ms, "cpu_cores": os.cpu_count(),
                                "neighbour_lists_identical_on_sample": bool(np.array_equal(g_sub[1], c_nn))}
    # rank 2: static-point selection of one keyframe cloud against the window, and the overlap ratio
    g_cl["x"] += 0.05
    pos = np.zeros(3, dtype=np.float32)
    g_ms, (sel, cnt) = gpu_ms(lambda: traj.selectStaticPoints(g_cl, pos, 0.3))
    flat = np.ascontiguousarray(np.stack([g_cl["x"], g_cl["y"], g_cl["z"], g_cl["w"], g_cl["nx"], g_cl["ny"], g_cl["nz"], g_cl["nw"]], 1).astype(np.float32))
    msq = np.float32(np.float64(np.float32(0.3)) ** 2)  # (float)std::pow(1.0f * minGridSize, 2), DmsaSlam.h:293
    c_ms, (sel_o, cnt_o) = cpu_ms(lambda: ob.select_static_points(world, flat, pos, msq, 0.3))
    out["select_static_points"] = {"window_points": int(len(world)), "keyframe_points": int(len(g_cl)), "selected": int(cnt), "gpu_ms_grid_cached": g_ms,
                                   "cpu_ms": c_ms, "cpu_cores": os.cpu_count(), "identical": bool(cnt == cnt_o and np.array_equal(sel, sel_o))}
    act = np.ascontiguousarray(np.stack([g_cl["x"], g_cl["y"], g_cl["z"], g_cl["w"]], 1)[sel.astype(bool)])
    g_ms, ov = gpu_ms(lambda: traj.overlap(act, 0.3))
    c_ms, ov_o = cpu_ms(lambda: ob.overlap(act, world, 0.3))
    out["overlap"] = {"map_points": int(len(act)), "gpu_ms": g_ms, "cpu_ms": c_ms, "value": ov, "identical": bool(ov == ov_o)}
    return out


def pinned_copy(arr):
    import torch

    t = torch.empty(arr.nbytes, dtype=torch.uint8, pin_memory=True)
    v = t.numpy().view(arr.dtype).reshape(arr.shape)
    v[...] = arr
    return t, v


def run_sliding(args):
    import torch
    import torch.distributed as dist

    from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    win = synth.make_config(args.config)  # every rank gets the same window: per-N values compare like with like
    torch.cuda.set_stream(torch.cuda.Stream(dev))  # a real (non-default) stream: the library launches on it, torch events see it
    stream = torch.cuda.current_stream().cuda_stream
    assert stream != 0
    s = DmsaOptimSettings(**SETTINGS)
    traj = ContinuousTrajectory.from_window(win, device=local, stream=stream)
    if args.lm_solver:
        traj.setLmSolver(args.lm_solver)
    traj.centralize()
    P = traj.numParams
    rel_o, rel_t = win["rel_orient"].copy(), win["rel_transl"].copy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # the window stays centralised between steps; resetting the poses re-applies the centralised pose 0
    pose0 = traj.getPoses()

    def reset_poses():
        traj.setRelativePoses(pose0["rel_orient"], pose0["rel_transl"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident-data throughput (value) ----------------
    last = None
    for _ in range(max(args.warmup, 3)):
        reset_poses()
        last = traj.iteration(s)
    G, M = traj.buildSets(s)  # sizes of the first iteration's sets (reported with every figure)
    sets = traj.getSets()
    C_FUSE = traj.L.dmsa_b200_fuse_threshold()
    N_points = int(traj.numPoints)
    reset_poses()
    traj.profileEnable(True)
    clocks = ClockSampler(local)
    clocks.start()
    l0 = traj.ctx.launch_count
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xff)  # L2 flush between timed steps (256 MiB write), outside the timed span
        reset_poses()
        ev[k][0].record()
        last = traj.iteration(s)
        ev[k][1].record()
    barrier()
    launches = traj.ctx.launch_count - l0
    prof = traj.profileRead()
    traj.profileEnable(False)
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps / (ms_total * 1e-3)

    # ---------------- end to end through the C-ABI with HOST buffers (e2e) ----------------
    pinned = [pinned_copy(sc) for sc in win["scans"]]
    pstat = pinned_copy(win["static"])
    h2d = sum(sc.nbytes for sc in win["scans"]) + win["static"].nbytes + 2 * rel_o.nbytes
    d2h = 4 * rel_o.nbytes + 64
    e2e_steps = max(3, min(args.steps, 10))

    from dmsa_lidar_slam_b200 import DmsaOptimizer

    def stage_window():
        traj.initTraj(win["t_min"], win["t_max"], win["n_poses"], False, win["dt_res"])
        traj.registerPcBuffer([p[1] for p in pinned], win["grid_sizes"])
        traj.addStaticPoints(pstat[1])
        traj.setRelativePoses(rel_o, rel_t)

    def e2e_step():  # one iteration per uploaded window (the conservative definition of round 1)
        stage_window()
        traj.centralize()
        d = traj.iteration(s)
        poses = traj.getPoses()  # result read-back
        return 1

    s_prod = DmsaOptimSettings(**dict(SETTINGS, num_iter=PRODUCTION_NUM_ITER))
    opt = DmsaOptimizer()

    def e2e_optimize():  # the call a user of the reference makes: optimizeSet on a freshly staged window (DmsaSlam.h:150-166)
        stage_window()
        rep = opt.optimizeSet(traj, s_prod)
        poses = traj.getPoses()
        return rep["iterations"]

    def time_e2e(fn, reps):
        for _ in range(2):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        t0 = time.perf_counter()
        iters = 0
        for k in range(reps):
            evs[k][0].record()
            iters += fn()
            evs[k][1].record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([max(ms, 1e3 * wall)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * iters / (float(t.item()) * 1e-3), iters / reps

    # the resident production loop: ONE optimizeSet call of `steps` loop bodies on the staged window (run-ahead loop: the host
    # consumes every body's read-back one body late; no L2 flush between bodies, which is how the loop runs in production)
    s_loop = DmsaOptimSettings(**dict(SETTINGS, num_iter=max(args.steps, 3), epsilon=0.0))
    loop = None
    try:
        for _ in range(2):
            reset_poses()
            opt.optimizeSet(traj, s_loop)
            traj.centralize()
        reset_poses()
        barrier()
        la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        la.record()
        rep_loop = opt.optimizeSet(traj, s_loop)
        lb.record()
        barrier()
        traj.centralize()
        lt = torch.tensor([la.elapsed_time(lb)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        loop = {"value": world * rep_loop["iterations"] / (float(lt.item()) * 1e-3), "unit": UNIT, "iterations": rep_loop["iterations"], "stop": rep_loop["stop"],
                "ms_per_iteration": float(lt.item()) / max(rep_loop["iterations"], 1),
                "what": "one optimizeSet call (centralize + loop bodies + decentralize + final updateGlobalPoints) on the resident window, run-ahead loop, no L2 flush between bodies"}
    except Exception as e:
        loop = {"error": repr(e)}
    reset_poses()
    e2e_single, _ = time_e2e(e2e_step, e2e_steps)
    e2e_value, e2e_iters = time_e2e(e2e_optimize, max(3, min(args.steps // 2, 6)))
    clk = clocks.stop()
    frontend = None
    if world == 1 and args.frontend:
        try:
            frontend = measure_frontend(traj, win, local)
        except Exception as e:  # never lose the headline over a side measurement
            frontend = {"error": repr(e)}
    keyframe = None
    if args.keyframe:
        from dmsa_lidar_slam_b200 import distributed

        del traj
        keyframe = distributed.measure_keyframe(max(3, min(args.steps, 10)), args.warmup, rank, world, local,
                                                cpu_baseline_fn=keyframe_cpu_baseline if world == 1 else None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    T = C_FUSE
    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = load_peaks()
    # the dominant single kernel: the fused evaluation of the forward-difference batch over the sets that fit one block (the
    # chunked pair k_cost_sum / k_cost_quad over the bigger sets is listed next to it in device_ms_per_step_breakdown)
    dom = "k_cost_fused_fd" if prof.get("k_cost_fused_fd", (0, 0))[1] else max((k for k in prof if k.startswith("k_cost")), key=lambda k: prof[k][0])
    dom_ms, dom_n = prof[dom]
    V = P + 1 if dom.endswith("_fd") else 9
    n_per_set = np.diff(sets["offs"])
    small = n_per_set <= T
    if "fused" in dom:
        # both passes of the sets it owns: SURVEY §8d Jacobian-pass figure, 24 B per membership + 48 B per set, once for all V vectors
        units_M, units_G = int(n_per_set[small].sum()), int(small.sum())
        alg_bytes = 24.0 * units_M + 48.0 * units_G
        flop_alg = 50.0 * units_M * V
    else:
        # one of the two passes over the big sets: one 16-byte member record per membership + the per-set record
        units_M, units_G = int(n_per_set[~small].sum()), int((~small).sum())
        alg_bytes = 16.0 * units_M + 48.0 * units_G
        flop_alg = 25.0 * units_M * V
    avg_ms = dom_ms / max(dom_n, 1)
    ach = alg_bytes / (avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tf = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tf) and args.config == "cfg2":
        tj = json.load(open(tf))
        if dom in tj:
            traffic, traffic_src = tj[dom], tj["_source"]
    breakdown = {k: round(v[0] / max(args.steps, 1), 4) for k, v in prof.items() if v[1]}
    # the kernel family that IS memory-shaped by SURVEY §8d: set construction (keys, sort, segmentation, statistics), 24 B per
    # point (read the world point, write the member record), 4 B per membership (the sorted index) and 48 B per set (its record)
    sets_ms = sum(v for k, v in breakdown.items() if k.startswith("sets_"))
    sets_bytes = 24.0 * N_points + 4.0 * float(M) + 48.0 * float(G)
    sets_roofline = {"phases": [k for k in breakdown if k.startswith("sets_")], "ms_per_build": sets_ms, "algorithmic_bytes_per_build": sets_bytes,
                     "achieved": sets_bytes / (sets_ms * 1e-3) / 1e9 if sets_ms > 0 else None, "unit": "GB/s",
                     "frac": sets_bytes / (sets_ms * 1e-3) / 1e9 / peak if sets_ms > 0 else None,
                     "note": "a chain of ~20 dependent launches over 1.4 M keys (four radix digit passes, two chained scans): bound by the latency of "
                             "the chain, not by bandwidth (DESIGN.md §5)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 point arithmetic / f64 pose chain, sums, J^T J (the reference's types)", "data": "synthetic",
        "config": {"workload": workload_string(args.config, win, P, world),
                   "N": N_points, "M": int(M), "G": int(G), "l2": "flushed between timed steps (256 MiB write, untimed)",
                   "settings": SETTINGS, "lm_solver": args.lm_solver},
        "point_jacobians_per_s": value * M,
        "membership_evals_per_s": value * M * (P + 10),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d / e2e_iters), "d2h_bytes_per_step": int(d2h / e2e_iters),
                "iterations_per_window": e2e_iters, "h2d_bytes_per_window": int(h2d), "d2h_bytes_per_window": int(d2h),
                "what": f"the reference's call on a new window, timed whole: traj_init (timing tables recomputed and uploaded: table reuse switched off) + "
                        f"register_scans + add_static_points (pinned host AoS PointStampId) + set poses + optimizeSet(num_iter = {PRODUCTION_NUM_ITER}, the "
                        "production setting of config/slam_settings.yaml:22: centralize, iterations, decentralize, final updateGlobalPoints) + pose read-back; "
                        "value = iterations run / time",
                "single_iteration_per_upload": {"value": e2e_single, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                                                "what": "round-1 definition: one iteration per uploaded window (upload + centralize + 1 iteration + pose read-back)"}},
        "gpu_launches": int(launches),
        "gpu_launches_note": "every kernel of the step is hand-written (the radix sort and the scans of the set build included: no library primitive on the path)",
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes, "memberships_per_launch": units_M, "sets_per_launch": units_G,
                     "vectors_per_launch": V, "peak_source": peak_src,
                     "fp32": {"algorithmic_flop_per_launch": flop_alg, "achieved_tflops": flop_alg / (avg_ms * 1e-3) / 1e12,
                              "note": "the forward-difference formulation is FP32-issue bound, not HBM bound (SURVEY §8d, DESIGN.md)"},
                     "set_build": sets_roofline,
                     # SURVEY §8d, whole loop body: B_alg = 24 N + 52 M + 144 G + 48 n_total (P + 10) + 8 (P^2 + P + 10) bytes, F_alg = 50 M (P + 10) + 2 G P^2 flop
                     "whole_iteration": (lambda B, F: {"algorithmic_bytes": B, "achieved": B * value / 1e9, "unit": "GB/s", "frac": B * value / 1e9 / peak,
                                                       "algorithmic_flop": F, "achieved_tflops": F * value / 1e12})(
                         24.0 * N_points + 52.0 * float(M) + 144.0 * float(G) + 48.0 * float(traj.timing()["n_total"]) * (P + 10) + 8.0 * (P * P + P + 10),
                         50.0 * float(M) * (P + 10) + 2.0 * float(G) * P * P)},
        "device_ms_per_step_breakdown": breakdown,
        "clocks": clk,
        "last_step": {"G": last["num_gaussians"], "error0": last["error0"], "best_step": last["best_step"], "stop": last["stop"]},
    }
    if loop is not None:
        line["resident_optimize_loop"] = loop
    if keyframe is not None:
        line["keyframe"] = keyframe
    if frontend is not None:
        line["frontend"] = frontend
    if world == 1:
        threads = best_thread_count(win)
        cb = oracle_cpu_baseline(win, threads)
        line["cpu_baseline"] = {"value": 1.0 / cb["seconds_per_iteration"], "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "median of 2 full iterations of the same workload; oracle/dmsa_oracle.cpp -O2 -fopenmp, faithful arithmetic"}
        # the reference's own build and threading: -O1 (CMakeLists.txt:15), serial cost loops (DmsaOptimizer.h:56-57 only feeds Eigen's GEMM)
        cf = oracle_cpu_baseline(win, 1, iters=1, opt="O1")
        line["cpu_baseline"]["reference_faithful"] = {"value": 1.0 / cf["seconds_per_iteration"], "unit": UNIT, "cores": 1, "kind": "port",
                                                      "sample": "1 full iteration; oracle/dmsa_oracle.cpp -O1, one thread (the reference's cost loops are serial)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sliding", choices=["sliding", "keyframe"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--lm-solver", type=int, default=0, help="dmsa_b200_set_lm_solver: 0 default (device LU for P <= 128, host beyond), 1 host, 2 device Cholesky")
    ap.add_argument("--frontend", type=int, default=1, help="1 (N = 1 only): also time SURVEY 8(f) ranks 2 / 3 (static points, pre-processing, normals) -> `frontend` object")
    ap.add_argument("--keyframe", type=int, default=1, help="1: also measure BASELINE config 4 (keyframe bundles, NCCL all-reduce) -> `keyframe` object")
    args = ap.parse_args()
    global RESULT_OUT
    RESULT_OUT = _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "keyframe":
        from dmsa_lidar_slam_b200 import distributed

        return distributed.bench_keyframe(args)
    return run_sliding(args)


if __name__ == "__main__":
    main()
