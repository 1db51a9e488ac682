"""Phase cycle counts of the device LM solve (debug): DMSA_B200_SOLVE_CLK=1 python scripts/solve_clk.py [n ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DMSA_B200_SOLVE_CLK"] = "1"
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings  # noqa: E402

t = ContinuousTrajectory()
for n in [int(a) for a in sys.argv[1:]] or [114, 234]:
    rng = np.random.default_rng(n)
    J = rng.standard_normal((3 * n, n))
    r = rng.standard_normal(3 * n)
    hg = np.concatenate([(J.T @ J).ravel(), J.T @ r, [float(r @ r)]])
    for _ in range(3):
        t.lmSolveDevice(DmsaOptimSettings(step_length_optim=0.2, max_step=0.3), hg, n)
