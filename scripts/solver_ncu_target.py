"""ncu target: a few device LM solves at n = 114 (scripts/solver_bench.py measures; this one is for --set full captures)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 114
traj = ContinuousTrajectory()
rng = np.random.default_rng(n)
J = rng.standard_normal((3 * n, n)) * np.logspace(0, -3, n)[None, :]
r = rng.standard_normal(3 * n)
hg = np.concatenate([(J.T @ J).ravel(), J.T @ r, [float(r @ r)]])
s = DmsaOptimSettings(step_length_optim=0.2, max_step=0.3, lambda_diag=1e-5)
for _ in range(4):
    traj.lmSolveDevice(s, hg, n)
