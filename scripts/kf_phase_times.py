"""Where does a keyframe-bundle iteration spend its time on ONE rank of an 8-rank job?  (rank 0's bundle alone on one GPU;
the all-reduces are skipped: emulate = True.)  Wall-clock per phase with a stream synchronisation after each."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from dmsa_lidar_slam_b200 import DmsaOptimSettings, synth
from dmsa_lidar_slam_b200.distributed import KEYFRAME_SETTINGS, KeyframeBundleOptimizer

sm = synth.make_keyframe_submap(n_keyframes=64, n_points=100000, seed=4)
s = DmsaOptimSettings(**KEYFRAME_SETTINGS)
o = KeyframeBundleOptimizer(sm, s, rank=0, world=8, emulate=True)
P = o.P
p0 = o.p.copy()
for _ in range(3):
    o.p = p0.copy()
    o.iteration()
acc = {}


def lap(name, t0):
    o.stream.synchronize()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + 1e3 * (t1 - t0)
    return t1


n = 10
tw0 = time.perf_counter()
for _ in range(n):
    o.p = p0.copy()
    o.iteration()
o.stream.synchronize()
whole = 1e3 * (time.perf_counter() - tw0) / n
for _ in range(n):
    o.p = p0.copy()
    with torch.cuda.stream(o.stream):
        t = time.perf_counter()
        o._push_poses()
        t = lap("push_poses (host: params -> relative -> global poses, set_relative_poses)", t)
        o.ghg.zero_(); o.out.zero_()
        for c, idx in zip(o.ctx, o.idx_dev):
            c.bundleJacobian(s, idx.data_ptr(), P, o.ghg.data_ptr(), False)
        t = lap("bundle_jacobian (tables, set build with splitSet, cost, J^T J, scatter)", t)
        o.c0.spdSolveDev(s, o.ghg.data_ptr(), P, o.out.data_ptr(), o.out.data_ptr() + 8 * P)
        t = lap("spd_solve_dev (P = 378 Cholesky LM step)", t)
        for c, idx in zip(o.ctx, o.idx_dev):
            c.bundleLineSearch(o.out.data_ptr(), idx.data_ptr(), o.out.data_ptr() + 8 * (P + 2))
        t = lap("bundle_line_search", t)
        o.host.copy_(o.out, non_blocking=True)
        o.stream.synchronize()
        for c in o.ctx:
            c.bundleVerify()
        h = o.host.numpy().copy()
        t = lap("read-back + bundle_verify", t)
print(json.dumps({"iteration_ms_one_bundle_rank": whole, "phases_ms": {k: round(v / n, 4) for k, v in acc.items()}}))
