"""One forward-difference cost+Jacobian evaluation on a BASELINE config: the command line profiled by ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings, synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
win = synth.make_config(cfg)
s = DmsaOptimSettings(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30)
t = ContinuousTrajectory.from_window(win)
t.centralize()
for _ in range(reps):
    d = t.iteration(s)
print(cfg, d["num_gaussians"], d["error0"])
