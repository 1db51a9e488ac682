"""Multi-GPU keyframe pass: keyframe bundles sharded across ranks, one all-reduce of [H | g | err0] and one of the 9
line-search costs per iteration.  One process per GPU, `torch.distributed` (NCCL over NVLink) for the
exchange; the sliding-window pass is single-GPU by design (north star) and never comes through here.

The reference optimises ONE keyframe submap (DmsaSlam.h:212-238, MapManagement.h:254-276).  A *bundle* is a contiguous
keyframe range == `getSubmap(from, to)`; because the parameters are RELATIVE poses (Poses.h:64-70), bundle-local
parameter i of keyframe from+j is global parameter of keyframe from+j, so a bundle's H_b / g_b scatter into the global
(block-banded) system.  With a single bundle covering all keyframes the scheme is exactly the reference's iteration.
This is a documented extension (BASELINE.json config 4), not reference behaviour.
"""
from __future__ import annotations

import os
import time

import numpy as np
from scipy.spatial.transform import Rotation as Rot


def bundle_ranges(n_keyframes, size, overlap):
    """Keyframe ranges [from, to] of `size` keyframes overlapping by `overlap` (cfg 4: 64 keyframes, 15 / 8 -> 8 bundles)."""
    size = min(size, n_keyframes)
    step = max(1, size - overlap)
    starts = list(range(0, max(1, n_keyframes - size + 1), step))
    if starts[-1] + size < n_keyframes:
        starts.append(n_keyframes - size)
    return [(s, s + size - 1) for s in starts]


def bundle_param_index(n_keyframes, first, last):
    """Global indices of a bundle's local parameter vector.

    Parameter layout (Poses.h:64-76): [w_1 .. w_{n-1} | t_1 .. t_{n-1}] — all orientations, then all translations,
    pose 0 excluded.  The bundle's pose 0 is keyframe `first` (held fixed inside the bundle)."""
    n, m = n_keyframes, last - first + 1
    idx = np.empty(6 * (m - 1), dtype=np.int64)
    for j in range(1, m):
        k = first + j  # global keyframe
        for a in range(3):
            idx[3 * (j - 1) + a] = 3 * (k - 1) + a
            idx[3 * (m - 1) + 3 * (j - 1) + a] = 3 * (n - 1) + 3 * (k - 1) + a
    return idx


def hg_scatter_index(idx, P):
    """Flat positions in the global [H (P*P row-major) | g (P) | err0] buffer of a bundle's [H_b | g_b | err0_b]."""
    Pb = len(idx)
    Hpos = (idx[:, None] * P + idx[None, :]).reshape(-1)
    return np.concatenate([Hpos, P * P + idx, [P * P + P]]).astype(np.int64), Pb


def relative2global(rel_o, rel_t):
    """ConsecutivePoses.h:26-43 on 3 x n arrays: the library's own host pose chain (a few microseconds for 64 keyframes)."""
    from .api import relative2global as r2g

    return r2g(rel_o, rel_t)


def relative2global_scipy(rel_o, rel_t):
    """The same chain through scipy (independent check of the library's, tests/test_distributed_cpu.py)."""
    n = rel_o.shape[1]
    E = Rot.from_rotvec(rel_o.T).as_matrix()  # exp of every relative orientation
    Rg = np.empty((n, 3, 3))
    gt = np.zeros_like(rel_t)
    R, T = np.eye(3), np.zeros(3)
    for k in range(n):
        T = T + R @ rel_t[:, k]
        gt[:, k] = T
        R = R @ E[k]
        Rg[k] = R
    go = Rot.from_matrix(Rg).as_rotvec().T
    return np.ascontiguousarray(go), gt


def params_to_rel(p, rel_o0, rel_t0, n):
    rel_o = np.zeros((3, n))
    rel_t = np.zeros((3, n))
    rel_o[:, 0], rel_t[:, 0] = rel_o0, rel_t0
    rel_o[:, 1:] = p[: 3 * (n - 1)].reshape(n - 1, 3).T
    rel_t[:, 1:] = p[3 * (n - 1):].reshape(n - 1, 3).T
    return rel_o, rel_t


def rel_to_params(rel_o, rel_t):
    return np.concatenate([rel_o[:, 1:].T.reshape(-1), rel_t[:, 1:].T.reshape(-1)])


class Exchange:
    """Sum over the ranks of a torch.distributed group (any backend: gloo CPU tensors in the CPU protocol test).  The GPU
    data path does NOT come through here: it uses the library's own NCCL communicator (dmsa_b200_comm_init)."""

    def __init__(self, world=1):
        self.world = world

    def all_reduce_sum(self, t):
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t


def select_step(error0, ls_costs):
    """adaptiveStepSize's arg-min with strict improvement (DmsaOptimizer.h:154-179)."""
    best, m = 0, error0
    for k in range(1, 10):
        if ls_costs[k - 1] < m:
            m, best = ls_costs[k - 1], k
    return best


def shared_comm_id(rank, world):
    """128-byte NCCL id of the library's own communicator: created on rank 0, handed to the other ranks through the
    (already initialised) torch.distributed group - rendezvous plumbing only, the data path is the library's NCCL calls."""
    from .api import comm_unique_id

    if world == 1:
        return None
    import torch.distributed as dist

    ids = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]


class KeyframeBundleOptimizer:
    """One keyframe-pass iteration over several GPUs, one process per GPU.

    * several bundles (BASELINE config 4): the bundles go round-robin over the ranks, one C-ABI context per bundle; an
      iteration is  bundle_jacobian (every local bundle: scatter-add of [H_b | g_b | err0_b | #sets] into the global buffer)
      -> all-reduce (P^2 + P + 2 doubles) -> device Cholesky LM step of the global system -> bundle_line_search (every local
      bundle) -> all-reduce (9 doubles) -> ONE read-back of [step | err0 | flag | 9 costs | #sets].  Every rank keeps the global
      parameter vector and takes the same step (identical all-reduced inputs, deterministic solver).
    * one bundle == the reference's keyframe pass (one submap, DmsaSlam.h:212-238): every rank stages the whole submap,
      owns the sets g % world == rank, and dmsa_b200_iteration does the two all-reduces itself (row sharding)."""

    def __init__(self, submap, settings, bundle_size=15, overlap=8, rank=0, world=1, device=0, stream=None, comm_id=None, emulate=False):
        """emulate = True: partition the bundles for (rank, world) but create no communicator (single-process tests that add the
        ranks' partial buffers themselves)."""
        import torch

        from .api import MapManagement

        self.torch = torch
        self.settings = settings
        self.rank, self.world = rank, world
        self.n = submap["n_keyframes"]
        self.P = 6 * (self.n - 1)
        self.rel_o0 = submap["rel_orient"][:, 0].copy()
        self.rel_t0 = submap["rel_transl"][:, 0].copy()
        self.p = rel_to_params(submap["rel_orient"], submap["rel_transl"])
        self.ranges = bundle_ranges(self.n, min(bundle_size, self.n), overlap)
        self.single = len(self.ranges) == 1
        self.dev = torch.device("cuda", device)
        # one CUDA stream for everything the contexts of this rank launch (Handle 0 == "create an own stream" in the C-ABI)
        self.stream = torch.cuda.Stream(self.dev) if not stream else torch.cuda.ExternalStream(stream, self.dev)
        stream = self.stream.cuda_stream
        self.emulate = emulate
        if comm_id is None and world > 1 and not emulate:
            comm_id = shared_comm_id(rank, world)
        self.ctx, self.idx, self.idx_dev = [], [], []
        if self.single:
            self.mine = list(self.ranges)
            c = MapManagement.from_submap(submap, device=device, stream=stream)
            if world > 1:
                c.commInit(comm_id, rank, world)  # row sharding: sets g % world == rank
            self.ctx.append(c)
            return
        self.mine = [b for i, b in enumerate(self.ranges) if i % world == rank]
        for (f, l) in self.mine:
            sm = dict(n_keyframes=l - f + 1, clouds=submap["clouds"][f:l + 1], rings=submap["rings"][f:l + 1], grid_sizes=submap["grid_sizes"][f:l + 1],
                      rel_orient=submap["rel_orient"][:, f:l + 1].copy(), rel_transl=submap["rel_transl"][:, f:l + 1].copy())
            self.ctx.append(MapManagement.from_submap(sm, device=device, stream=stream))
            idx = bundle_param_index(self.n, f, l)
            self.idx.append(idx)
            self.idx_dev.append(torch.from_numpy(idx.astype(np.int32)).to(self.dev))
        # the communicator lives in one context of the rank (a rank without bundles still takes part in the collectives)
        self.c0 = self.ctx[0] if self.ctx else MapManagement(2, device=device, stream=stream)
        if world > 1 and not emulate:
            self.c0.commInit(comm_id, rank, world)
            self.c0.setShard(0, 1)  # bundles are whole on their rank: no row sharding inside a bundle
        P = self.P
        self.ghg = torch.zeros(P * P + P + 3, dtype=torch.float64, device=self.dev)  # [H | g | err0 | number of sets | missed guesses]
        self.out = torch.zeros(P + 2 + 9 + 2, dtype=torch.float64, device=self.dev)  # [step | err0, flag | 9 costs | number of sets, misses]
        self.host = torch.zeros(P + 2 + 9 + 2, dtype=torch.float64).pin_memory()
        self.num_sets = 0

    def _push_poses(self):
        rel_o, rel_t = params_to_rel(self.p, self.rel_o0, self.rel_t0, self.n)
        if self.single:
            self.ctx[0].setRelativePoses(rel_o, rel_t)
            return
        go, gt = relative2global(rel_o, rel_t)
        for c, (f, l) in zip(self.ctx, self.mine):
            ro, rt = rel_o[:, f:l + 1].copy(), rel_t[:, f:l + 1].copy()
            ro[:, 0], rt[:, 0] = go[:, f], gt[:, f]  # the bundle's pose 0 is keyframe f's current global pose
            c.setRelativePoses(ro, rt)

    def iteration(self):
        """One DMSA iteration.  Returns dict(stop, error0, best_step, step_norm, num_sets, ...)."""
        with self.torch.cuda.stream(self.stream):
            if self.single:
                return self._iteration_single()
            d = self._iteration_bundles(False)
            if d is None:  # a deferred set build ran on a wrong guess (rare): repeat with synchronous builds
                d = self._iteration_bundles(True)
            return d

    def _iteration_single(self):
        c = self.ctx[0]
        self._push_poses()
        d = c.iteration(self.settings)
        self.p = c.getPoseParameters()
        self.num_sets = d["num_gaussians"]
        return dict(stop=d["stop"], error0=d["error0"], best_step=d["best_step"], step_norm=d["step_norm"], num_sets=d["num_gaussians"],
                    ls=d["ls_cost"], step=d["step"])

    def jacobian_phase(self, sync_build=False):
        """This rank's partial [H | g | err0 | #sets | #misses] in self.ghg (before the all-reduce).  Everything is enqueued on
        self.stream (torch's zero_ included); a caller outside iteration() synchronises before reading self.ghg."""
        s, P = self.settings, self.P
        with self.torch.cuda.stream(self.stream):
            self._push_poses()
            self.ghg.zero_()
            self.out.zero_()
            for c, idx in zip(self.ctx, self.idx_dev):
                c.bundleJacobian(s, idx.data_ptr(), P, self.ghg.data_ptr(), sync_build)

    def _iteration_bundles(self, sync_build):
        s, P = self.settings, self.P
        self.jacobian_phase(sync_build)
        ghg, out = self.ghg.data_ptr(), self.out.data_ptr()
        self.c0.allReduce(ghg, P * P + P + 3)  # exchange 1: [H | g | err0 | #sets | #missed guesses]
        self.c0.spdSolveDev(s, ghg, P, out, out + 8 * P)
        for c, idx in zip(self.ctx, self.idx_dev):
            c.bundleLineSearch(out, idx.data_ptr(), out + 8 * (P + 2))
        self.c0.allReduce(out + 8 * (P + 2), 9)  # exchange 2: 9 trial costs
        self.out[P + 11:P + 13].copy_(self.ghg[P * P + P + 1:P * P + P + 3])
        self.host.copy_(self.out, non_blocking=True)
        self.stream.synchronize()  # the iteration's only host wait
        for c in self.ctx:
            c.bundleVerify()  # refreshes the contexts' guesses for the next build
        h = self.host.numpy()
        if h[P + 12] != 0 and not sync_build:  # some bundle on some rank missed its guess: every rank sees the same count
            return None
        step, error0, flag = h[:P].copy(), float(h[P]), int(h[P + 1])
        ls = h[P + 2:P + 11].copy()
        self.num_sets = int(round(h[P + 11]))
        if self.num_sets < s.min_num_gaussians:  # DmsaOptimizer.h:89-93
            return dict(stop="few_gaussians", error0=0.0, best_step=0, step_norm=0.0, num_sets=self.num_sets)
        if flag == 2:  # not numerically positive definite: host LU solve of the same system, then the line search again
            from .api import lm_solve

            step, nan = lm_solve(s, self.ghg.cpu().numpy()[:P * P + P + 1], P, explicit_inverse=0)
            flag = 1 if nan else 0
            if not nan:
                self.out[:P].copy_(self.torch.from_numpy(step).to(self.dev))
                self.out[P + 2:P + 11].zero_()
                for c, idx in zip(self.ctx, self.idx_dev):
                    c.bundleLineSearch(out, idx.data_ptr(), out + 8 * (P + 2))
                self.c0.allReduce(out + 8 * (P + 2), 9)
                ls = self.out[P + 2:P + 11].cpu().numpy()
        if flag == 1:
            return dict(stop="nan", error0=error0, best_step=0, step_norm=0.0, num_sets=self.num_sets)
        best = select_step(error0, ls)
        nrm = float(np.linalg.norm(step))
        if best == 0:
            self.p = self.p + 0.9 * step  # DmsaOptimizer.h:130-134: no restore
            return dict(stop="no_improvement", error0=error0, best_step=0, step_norm=nrm, num_sets=self.num_sets, ls=ls, step=step)
        self.p = self.p + 0.1 * best * step
        stop = "epsilon" if nrm < s.epsilon else "max_iter"
        return dict(stop=stop, error0=error0, best_step=best, step_norm=nrm, num_sets=self.num_sets, ls=ls, step=step)

    def launch_count(self):
        return sum(c.ctx.launch_count for c in self.ctx)

    def collective_count(self):
        return int(self.ctx[0].collective_count) if self.single else int(self.c0.collective_count)


KEYFRAME_SETTINGS = dict(num_iter=1, step_length_optim=0.2, max_step=0.01, min_num_points_per_set=10, min_num_gaussians=30, gauss_split=1, epsilon=1e-4)


def measure_keyframe(steps, warmup, rank, world, local, n_kf=64, n_pts=100000, with_single_gpu=True, cpu_baseline_fn=None):
    """BASELINE config 4 on `world` ranks: 64 keyframes x 100k points, 8 overlapping bundles (15 keyframes, overlap 8) round-robin
    over the ranks, production keyframe settings (gauss_split = true, DmsaSlam.h:93), per iteration one NCCL all-reduce of
    [J^T J | J^T r | e^T e | #sets | #misses] and one of the 9 line-search costs.  Strong scaling: the total work is fixed.
    Also measures, in the same job, the same problem on rank 0's GPU alone (the strong-scaling denominator).
    torch.distributed must be initialised by the caller when world > 1.  Returns the `keyframe` object of the bench line (rank 0)."""
    import torch
    import torch.distributed as dist

    from . import synth
    from .api import DmsaOptimSettings

    sm = synth.make_keyframe_submap(n_keyframes=n_kf, n_points=n_pts, seed=4)
    s = DmsaOptimSettings(**KEYFRAME_SETTINGS)
    warmup = max(warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(opt, nranks):
        torch.cuda.set_stream(opt.stream)  # events below are recorded on the stream the kernels run on
        p0 = opt.p.copy()
        last = None
        for _ in range(warmup):
            opt.p = p0.copy()
            last = opt.iteration()
        l0, c0 = opt.launch_count(), opt.collective_count()
        if nranks > 1:
            barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            opt.p = p0.copy()
            last = opt.iteration()
        e1.record()
        if nranks > 1:
            barrier()
        else:
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if nranks > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=opt.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # the two exchanges alone, on the same buffers (device time of the collectives of one iteration)
        ar_us = None
        if nranks > 1 and not opt.single:
            P = opt.P
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a0.record()
            for _ in range(20):
                opt.c0.allReduce(opt.ghg.data_ptr(), P * P + P + 3)
                opt.c0.allReduce(opt.out.data_ptr() + 8 * (P + 2), 9)
            a1.record()
            torch.cuda.synchronize()
            ar_us = 1e3 * a0.elapsed_time(a1) / 20
        return dict(ms_per_iteration=ms / steps, iterations_per_s=steps / (ms * 1e-3), launches=int(opt.launch_count() - l0),
                    collectives=int(opt.collective_count() - c0), all_reduce_us_per_iteration=ar_us, last=last)

    opt = KeyframeBundleOptimizer(sm, s, 15, 8, rank, world, local, None)
    multi = timed(opt, world)
    P, nb = opt.P, len(opt.ranges)
    err_multi = multi["last"]["error0"]
    del opt
    single = None
    if world > 1 and with_single_gpu:
        if rank == 0:
            o1 = KeyframeBundleOptimizer(sm, s, 15, 8, 0, 1, local, None)
            single = timed(o1, 1)
            del o1
        barrier()
    elif world == 1:
        single = multi
    if rank != 0:
        return None
    kf = {
        "workload": f"cfg4: keyframe-graph DMSA iteration, {n_kf} keyframes x {n_pts} pts, {nb} overlapping bundles (15 keyframes, overlap 8) round-robin over "
                    f"{world} rank(s); per iteration one NCCL all-reduce of [J^T J | J^T r | e^T e | #sets | #misses] and one of the 9 line-search costs "
                    "(library-owned communicator), device Cholesky LM step of the global system, one read-back",
        "settings": KEYFRAME_SETTINGS, "P": P, "n_gpus": world, "scaling": "strong",
        "iterations_per_s": multi["iterations_per_s"], "ms_per_iteration": multi["ms_per_iteration"],
        "iterations_per_s_1gpu_same_job": single["iterations_per_s"] if single else None,
        "speedup_vs_1gpu": (multi["iterations_per_s"] / single["iterations_per_s"]) if single else None,
        "strong_scaling_efficiency": (multi["iterations_per_s"] / single["iterations_per_s"] / world) if single else None,
        "all_reduce_bytes_per_iteration": 8 * (P * P + P + 3) + 72, "all_reduce_us_per_iteration": multi["all_reduce_us_per_iteration"],
        "collectives_per_iteration": multi["collectives"] / max(steps, 1), "gpu_launches_rank0": multi["launches"],
        "sets": multi["last"]["num_sets"], "error0": err_multi, "error0_1gpu": single["last"]["error0"] if single else None,
        "best_step": multi["last"]["best_step"], "stop": multi["last"]["stop"],
    }
    if cpu_baseline_fn is not None:  # bench.py's CPU leg (the product package itself never touches the oracle)
        kf["cpu_baseline"] = cpu_baseline_fn(sm, nb, KEYFRAME_SETTINGS)
    return kf


def bench_keyframe(args):
    """`bench.py --workload keyframe`: only the config-4 keyframe pass (the default bench line carries it as its `keyframe` object)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __main__ as _m

    kf = measure_keyframe(args.steps, args.warmup, rank, world, local, int(os.environ.get("DMSA_KF", "64")), int(os.environ.get("DMSA_KF_POINTS", "100000")),
                          cpu_baseline_fn=getattr(_m, "keyframe_cpu_baseline", None) if world == 1 else None)
    if rank == 0:
        line = {"metric": "DMSA iterations/sec", "value": kf["iterations_per_s"], "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": kf["ms_per_iteration"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 point arithmetic / f64 pose chain, sums, J^T J", "data": "synthetic", "config": {"workload": kf["workload"]},
                "gpu_launches": kf["gpu_launches_rank0"], "keyframe": kf, "cpu_baseline": kf.get("cpu_baseline")}
        import json

        (_m.emit if hasattr(_m, "emit") else (lambda l: print(json.dumps(l))))(line)
    if world > 1:
        dist.destroy_process_group()
