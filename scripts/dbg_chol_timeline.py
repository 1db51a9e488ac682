"""Development aid: phase stamps of k_chol_solve (library built with DMSA_TIMELINE=3)."""
import ctypes, sys
sys.path.insert(0, ".")
import numpy as np
from dmsa_lidar_slam_b200 import DmsaOptimSettings, api
from dmsa_lidar_slam_b200.api import _Context, OptimizablePointSet
n = int(sys.argv[1]) if len(sys.argv) > 1 else 378
rng = np.random.default_rng(0)
A = rng.normal(size=(n + 50, n)); H = A.T @ A; g = rng.normal(size=n)
hg = np.concatenate([H.ravel(), g, [1.0]])
from dmsa_lidar_slam_b200 import MapManagement
m = MapManagement(2)
s = DmsaOptimSettings()
for _ in range(3):
    step, flag = m.spdSolve(s, hg, n)
L = api.load_library()
out = np.zeros(16384, dtype=np.uint64)
L.dmsa_b200_dbg_times(out.ctypes.data_as(ctypes.c_void_p))
a = out.reshape(-1, 8)[0].astype(np.int64)
print("block 0 stamps (ns from start): load", a[1] - a[0], "| diag(0)", a[2] - a[1], "sync", a[3] - a[2], "| panel(0)", a[4] - a[3], "sync+", a[5] - a[4], "(incl. update 0) | whole factor loop", a[6] - a[1], "| back substitution", a[7] - a[6], "| total", a[7] - a[0])
