"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
two iterations of the sliding-window pass on the tiny and cfg1 windows (pair-packed and scalar cost kernels, host and
device LM step, reference-order mean mode) and one keyframe iteration with gauss_split."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from dmsa_lidar_slam_b200 import (ContinuousTrajectory, DmsaOptimizer, DmsaOptimSettings, MapManagement, PreProcessor, PreprocessConfig,  # noqa: E402
                                  decode_pointcloud2, pc2_layout_for_sensor, synth)
from dmsa_lidar_slam_b200.synth import POINT_NORMAL  # noqa: E402

names = sys.argv[1:] or ["tiny"]
for name in names:
    win = synth.make_config(name)
    mp = 6 if name == "tiny" else 10
    s = DmsaOptimSettings(num_iter=2, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=mp, min_num_gaussians=10)
    for pair, solver, mean in ((1, 1, 0), (0, 0, 0), (1, 0, 1)):
        t = ContinuousTrajectory.from_window(win)
        t.setPairMode(pair)
        t.setLmSolver(solver)
        t.setMeanMode(mean)
        t.centralize()
        d = None
        for _ in range(2):
            d = t.iteration(s)
        print(name, "pair", pair, "device-solver", 1 - solver, "mean-mode", mean, d["num_gaussians"], d["error0"], flush=True)
try:
    sub = synth.make_keyframe_submap(n_keyframes=4, n_points=3000, seed=3)
    m = MapManagement.from_submap(sub)
    sk = DmsaOptimSettings(num_iter=1, step_length_optim=0.2, max_step=0.01, min_num_points_per_set=10, min_num_gaussians=5, gauss_split=True)
    d = m.iteration(sk)
    print("keyframes", d["num_gaussians"], d["error0"], flush=True)
except Exception as e:  # the keyframe generator's signature is not part of this script's contract
    print("keyframe case skipped:", e, flush=True)

# run-ahead optimize loop, pre-processing, normals, static-point selection, PointCloud2 decode (round 2 kernels)
win = synth.make_config("tiny")
t = ContinuousTrajectory.from_window(win)
rep = DmsaOptimizer().optimizeSet(t, DmsaOptimSettings(num_iter=4, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=6, min_num_gaussians=10))
print("run-ahead optimize", rep["iterations"], rep["stop"], flush=True)
pre = PreProcessor(ctx=t.ctx)
raw = win["scans"][0]
out, gs = pre.preProcess(raw, PreprocessConfig(500, 5.0, 0.1), 3)
filt, idx = pre.randomGridDownsampling(raw, 0.3, 3)
c = np.zeros(len(filt), dtype=POINT_NORMAL)
c["x"], c["y"], c["z"], c["w"] = filt["x"], filt["y"], filt["z"], 1.0
c, nn = pre.updateNormals(c, (0, 0, 0), 0.3, with_neighbours=True)
t.updateGlobalPoints()
sel, cnt = t.selectStaticPoints(c, np.zeros(3, dtype=np.float32), 0.3)
ov = t.overlap(np.stack([c["x"], c["y"], c["z"], c["w"]], 1), 0.3)
msg = np.zeros(1000, dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4"), ("ring", "<u2"), ("time", "<f4")]))
L = pc2_layout_for_sensor("velodyne", [0, 4, 8, 12, 16, 18], msg.dtype.itemsize)
dec = decode_pointcloud2(t.ctx, msg.tobytes(), len(msg), L, 1.0)
print("frontend", len(out), gs, len(idx), int(cnt), ov, len(dec), flush=True)
