"""Wall time of every C-ABI call of one end-to-end step (host buffers -> iteration -> poses), cfg2 by default."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
win = synth.make_config(cfg)
s = DmsaOptimSettings(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30)
traj = ContinuousTrajectory.from_window(win)


def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t, t.numpy().view(a.dtype).reshape(a.shape)


pinned = [pin(sc) for sc in win["scans"]]
pstat = pin(win["static"])
acc = {}


def timed(name, fn, *a):
    t0 = time.perf_counter()
    r = fn(*a)
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    return r


N = 12
for it in range(N + 2):
    if it == 2:
        acc.clear()
    timed("initTraj", traj.initTraj, win["t_min"], win["t_max"], win["n_poses"], False, win["dt_res"])
    timed("registerPcBuffer", traj.registerPcBuffer, [p[1] for p in pinned], win["grid_sizes"])
    timed("addStaticPoints", traj.addStaticPoints, pstat[1])
    timed("setRelativePoses", traj.setRelativePoses, win["rel_orient"], win["rel_transl"])
    timed("centralize", traj.centralize)
    timed("iteration", traj.iteration, s)
    timed("getPoses", traj.getPoses)
tot = sum(acc.values())
for k, v in acc.items():
    print(f"{k:18s} {1e3 * v / N:8.3f} ms")
print(f"{'total':18s} {1e3 * tot / N:8.3f} ms  ({N / tot:.1f} it/s)")
