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
    """ConsecutivePoses.h:26-43 on 3 x n arrays (host, double); rotations converted in two batched scipy calls."""
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
    """The two collectives of an iteration.  `group is None` and world 1 -> no-op; works on any torch backend
    (NCCL device tensors on the GPUs, gloo CPU tensors in the CPU tests)."""

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


class KeyframeBundleOptimizer:
    """Bundles of a keyframe submap, round-robin over ranks; every rank keeps the global parameter vector."""

    def __init__(self, submap, settings, bundle_size=15, overlap=8, rank=0, world=1, device=0, stream=None):
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
        self.mine = [b for i, b in enumerate(self.ranges) if i % world == rank]
        self.dev = torch.device("cuda", device)
        # one CUDA stream for everything: the library's kernels (contexts are created on it) and torch's index_add_/all_reduce,
        # so that the partial buffers are ordered without host synchronisation.  (Handle 0 == "create an own stream" in the C-ABI.)
        self.stream = torch.cuda.Stream(self.dev) if not stream else torch.cuda.ExternalStream(stream, self.dev)
        stream = self.stream.cuda_stream
        self.ex = Exchange(world)
        self.ctx, self.idx, self.scatter = [], [], []
        for (f, l) in self.mine:
            sm = dict(n_keyframes=l - f + 1, clouds=submap["clouds"][f:l + 1], rings=submap["rings"][f:l + 1], grid_sizes=submap["grid_sizes"][f:l + 1],
                      rel_orient=submap["rel_orient"][:, f:l + 1].copy(), rel_transl=submap["rel_transl"][:, f:l + 1].copy())
            self.ctx.append(MapManagement.from_submap(sm, device=device, stream=stream))
            idx = bundle_param_index(self.n, f, l)
            self.idx.append(idx)
            pos, _ = hg_scatter_index(idx, self.P)
            self.scatter.append(torch.from_numpy(pos).to(self.dev))
        self.hg = torch.zeros(self.P * self.P + self.P + 2, dtype=torch.float64, device=self.dev)  # [H | g | err0 | number of sets]
        self.ls = torch.zeros(9, dtype=torch.float64, device=self.dev)
        Pb = max([len(i) for i in self.idx] + [1])
        self.tmp = torch.zeros(Pb * Pb + Pb + 1, dtype=torch.float64, device=self.dev)
        self.tmp9 = torch.zeros(9, dtype=torch.float64, device=self.dev)
        self.num_sets = 0

    def _push_poses(self):
        rel_o, rel_t = params_to_rel(self.p, self.rel_o0, self.rel_t0, self.n)
        go, gt = relative2global(rel_o, rel_t)
        for c, (f, l) in zip(self.ctx, self.mine):
            ro, rt = rel_o[:, f:l + 1].copy(), rel_t[:, f:l + 1].copy()
            ro[:, 0], rt[:, 0] = go[:, f], gt[:, f]  # the bundle's pose 0 is keyframe f's current global pose
            c.setRelativePoses(ro, rt)

    def iteration(self):
        """One DMSA iteration over all bundles.  Returns dict(stop, error0, best_step, step_norm, num_sets)."""
        with self.torch.cuda.stream(self.stream):
            return self._iteration()

    def _iteration(self):
        torch = self.torch
        s = self.settings
        self._push_poses()
        self.hg.zero_()
        G = 0
        for c, pos, idx in zip(self.ctx, self.scatter, self.idx):
            c.updateGlobalPoints()
            g, _ = c.buildSets(s)
            G += g
            n = len(idx) * len(idx) + len(idx) + 1
            c.costJacobianDev(self.tmp.data_ptr())
            self.hg.index_add_(0, pos, self.tmp[:n])
        self.hg[-1] = float(G)
        self.ex.all_reduce_sum(self.hg)  # exchange 1: P*P + P + 2 doubles ([H | g | err0 | set count])
        P = self.P
        if len(self.ranges) == 1:
            # a single bundle is the reference's iteration: keep its arithmetic (host LU inverse), bit-equal to dmsa_b200_iteration
            hg = self.hg.cpu().numpy()
            error0, self.num_sets = float(hg[-2]), int(hg[-1])
            from .api import lm_solve

            step, nan = lm_solve(s, hg, P, explicit_inverse=1)
        else:
            # bundle extension (no reference arithmetic to mirror): solve the SPD system where it already lives — on the
            # device, with the library Cholesky behind torch.linalg (cuSOLVER; a plain library factorisation) — and bring back
            # only the step.  Every rank solves redundantly; identical inputs give identical steps (no broadcast needed).
            H = self.hg[:P * P].view(P, P).clone()
            H.diagonal().add_(float(np.float32(s.lambda_diag)))
            L, info = torch.linalg.cholesky_ex(H)
            x = torch.cholesky_solve(self.hg[P * P:P * P + P].unsqueeze(1), L).squeeze(1)
            st_dev = (-s.step_length_optim) * x
            m = st_dev.abs().max()
            st_dev = torch.where(m > s.max_step, st_dev * (s.max_step / m), st_dev)  # infinity-norm clamp, DmsaOptimizer.h:125-128
            out = torch.cat([st_dev, self.hg[-2:], info.to(torch.float64).view(1)]).cpu().numpy()
            step, error0, self.num_sets = out[:P].copy(), float(out[P]), int(out[P + 1])
            nan = bool(np.isnan(step).any()) or out[P + 2] != 0
            if out[P + 2] != 0:  # not numerically SPD: fall back to the host LU solve
                from .api import lm_solve

                step, nan = lm_solve(s, self.hg.cpu().numpy(), P, explicit_inverse=0)
        if self.num_sets < s.min_num_gaussians:  # DmsaOptimizer.h:89-93
            return dict(stop="few_gaussians", error0=0.0, best_step=0, step_norm=0.0, num_sets=self.num_sets)
        if nan:
            return dict(stop="nan", error0=error0, best_step=0, step_norm=0.0, num_sets=self.num_sets)
        self.ls.zero_()
        for c, idx in zip(self.ctx, self.idx):
            c.lineSearchCostsDev(step[idx], self.tmp9.data_ptr())
            self.ls += self.tmp9
        self.ex.all_reduce_sum(self.ls)  # exchange 2: 9 doubles
        ls = self.ls.cpu().numpy()
        best = select_step(error0, ls)
        nrm = float(np.linalg.norm(step))
        if best == 0:
            self.p = self.p + 0.9 * step  # DmsaOptimizer.h:130-134: no restore
            return dict(stop="no_improvement", error0=error0, best_step=0, step_norm=nrm, num_sets=self.num_sets, ls=ls)
        self.p = self.p + 0.1 * best * step
        stop = "epsilon" if nrm < s.epsilon else "max_iter"
        return dict(stop=stop, error0=error0, best_step=best, step_norm=nrm, num_sets=self.num_sets, ls=ls)

    def launch_count(self):
        return sum(c.ctx.launch_count for c in self.ctx)


def bench_keyframe(args):
    """`bench.py --workload keyframe`: BASELINE config 4 — 64 keyframes x 100k points, 8 overlapping bundles sharded over
    the ranks, NCCL all-reduce of J^T J / J^T r per iteration.  Strong scaling: the total work is fixed."""
    import json

    import torch
    import torch.distributed as dist

    from . import synth
    from .api import DmsaOptimSettings

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_kf = int(os.environ.get("DMSA_KF", "64"))
    n_pts = int(os.environ.get("DMSA_KF_POINTS", "100000"))
    sm = synth.make_keyframe_submap(n_keyframes=n_kf, n_points=n_pts, seed=4)
    st = dict(num_iter=1, step_length_optim=0.2, max_step=0.01, min_num_points_per_set=10, min_num_gaussians=30, gauss_split=0, epsilon=1e-4)
    s = DmsaOptimSettings(**st)
    opt = KeyframeBundleOptimizer(sm, s, 15, 8, rank, world, local, None)
    torch.cuda.set_stream(opt.stream)  # events / barriers below are recorded on the stream the kernels run on
    p0 = opt.p.copy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        opt.p = p0.copy()
        last = opt.iteration()
    l0 = opt.launch_count()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        opt.p = p0.copy()
        last = opt.iteration()
    e1.record()
    barrier()
    ms = max(e0.elapsed_time(e1), 0.0)
    t = torch.tensor([ms], dtype=torch.float64, device=opt.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = opt.launch_count() - l0
    if rank == 0:
        value = args.steps / (ms * 1e-3)
        line = {
            "metric": "DMSA iterations/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 point arithmetic / f64 pose chain, sums, J^T J", "data": "synthetic",
            "config": {"workload": f"cfg4: keyframe-graph DMSA iteration, {n_kf} keyframes x {n_pts} pts, {len(opt.ranges)} overlapping bundles (15 keyframes, overlap 8) "
                                   f"round-robin over {world} rank(s), NCCL all-reduce of [J^T J | J^T r | e^T e | #sets] ({opt.P * opt.P + opt.P + 2} doubles) + 9 line-search costs per iteration",
                       "P": opt.P, "sets": last["num_sets"], "settings": st, "l2": "working set (>= 8 x 1.5M points x 64 B) exceeds L2"},
            "gpu_launches": int(launches), "last_step": {k: (v if not isinstance(v, np.ndarray) else v.tolist()) for k, v in last.items()},
        }
        import __main__ as _m

        (_m.emit if hasattr(_m, "emit") else (lambda l: print(json.dumps(l))))(line)
    if world > 1:
        dist.destroy_process_group()
