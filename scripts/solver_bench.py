"""Device vs host LM step (DmsaOptimizer.h:107-128) at the BASELINE parameter counts: device time from CUDA events around the
three solver kernels (k_lu128 / k_inv128 / k_step_fin), host time as wall clock of dmsa_b200_lm_solve."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings  # noqa: E402
from dmsa_lidar_slam_b200.api import lm_solve  # noqa: E402

traj = ContinuousTrajectory()
out = {}
for n in (18, 84, 114, 128):
    rng = np.random.default_rng(n)
    J = rng.standard_normal((3 * n, n)) * np.logspace(0, -3, n)[None, :]
    r = rng.standard_normal(3 * n)
    hg = np.concatenate([(J.T @ J).ravel(), J.T @ r, [float(r @ r)]])
    s = DmsaOptimSettings(step_length_optim=0.2, max_step=0.3, lambda_diag=1e-5)
    for _ in range(3):
        traj.lmSolveDevice(s, hg, n)
    traj.profileEnable(True)
    reps = 50
    for _ in range(reps):
        b, _ = traj.lmSolveDevice(s, hg, n)
    ms, cnt = traj.profileRead()["k_lm_solve"]
    traj.profileEnable(False)
    res = {}
    for mode in (1, 2):
        for _ in range(3):
            a, _ = lm_solve(s, hg, n, mode)
        t0 = time.perf_counter()
        for _ in range(reps):
            a, _ = lm_solve(s, hg, n, mode)
        res[mode] = (time.perf_counter() - t0) / reps * 1e3
    out[n] = dict(device_ms=ms / cnt, host_serial_ms=res[1], host_helpers_ms=res[2], bit_identical=bool(np.array_equal(a, b)))
print(json.dumps(out, indent=1))
