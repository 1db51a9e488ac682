"""Timing of the SURVEY §8(f) rank-2 entry points at BASELINE config-2 scale: window cloud of cfg2 in HBM, one keyframe cloud
of 100 000 points (DmsaSlam.h:304-339) and the overlap ratio of the selected points (DmsaSlam.h:377-414); the oracle's grid
restatement on the host cores beside it.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from dmsa_lidar_slam_b200 import ContinuousTrajectory, synth  # noqa: E402

win = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
traj = ContinuousTrajectory.from_window(win)
traj.updateGlobalPoints()
W = np.ascontiguousarray(traj.globalPoints(), dtype=np.float32).reshape(-1, 4)
radius = np.float32(0.3)
rng = np.random.default_rng(3)
n = 100000
src = W[rng.integers(0, len(W), n), :3]
d = rng.normal(size=(n, 3))
d /= np.linalg.norm(d, axis=1, keepdims=True)
xyz = (src + d * rng.uniform(0, 0.8, n)[:, None]).astype(np.float32)
nrm = rng.normal(size=(n, 3))
nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
cloud = np.zeros(n, dtype=synth.POINT_NORMAL)
cloud["x"], cloud["y"], cloud["z"], cloud["w"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], 1.0
cloud["nx"], cloud["ny"], cloud["nz"] = nrm[:, 0].astype(np.float32), nrm[:, 1].astype(np.float32), nrm[:, 2].astype(np.float32)
pos = np.array([2.0, 15.0, 1.5], dtype=np.float32)
flat = np.zeros((n, 8), dtype=np.float32)
flat[:, :3], flat[:, 3], flat[:, 4:7] = xyz, 1.0, nrm.astype(np.float32)


def med(fn, reps=5):
    ts = []
    out = None
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts)), out


traj.selectStaticPoints(cloud, pos, radius)
t_sel, (sel, cnt) = med(lambda: traj.selectStaticPoints(cloud, pos, radius))
active = np.ones((int(sel.sum()), 4), dtype=np.float32)
active[:, :3] = xyz[sel == 1]
traj.overlap(active, radius)
t_ov, ov = med(lambda: traj.overlap(active, radius))
t_sel_cpu, (sel_o, cnt_o) = med(lambda: ob.select_static_points(W, flat, pos, np.float32(float(radius) ** 2), radius), reps=2)
t_ov_cpu, ov_o = med(lambda: ob.overlap(active, W, radius), reps=2)
print(json.dumps({"window_points": int(len(W)), "keyframe_points": n, "selected": int(cnt), "overlap": ov,
                  "bit_exact": bool(np.array_equal(sel, sel_o) and cnt == cnt_o and ov == ov_o),
                  "select_ms_gpu_host_buffers": 1e3 * t_sel, "overlap_ms_gpu_host_buffers": 1e3 * t_ov,
                  "select_ms_cpu_oracle": 1e3 * t_sel_cpu, "overlap_ms_cpu_oracle": 1e3 * t_ov_cpu, "cpu_threads": os.cpu_count()}))
