"""Development aid: phase time stamps (globaltimer) of the last launch of the instrumented kernel (library built with -DDMSA_TIMELINE)."""
import sys, ctypes
sys.path.insert(0, ".")
import numpy as np
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings, synth, api
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 346
win = synth.make_config("cfg2")
s = DmsaOptimSettings(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30)
t = ContinuousTrajectory.from_window(win)
t.centralize()
for _ in range(3):
    d = t.iteration(s)
t.ctx.synchronize()
L = api.load_library()
out = np.zeros(16384, dtype=np.uint64)
L.dmsa_b200_dbg_times(out.ctypes.data_as(ctypes.c_void_p))
a = out.reshape(-1, 8)[:nblk, :7].astype(np.int64)
t0 = a[:, 0].min()
a -= t0
np.set_printoptions(linewidth=200)
print("per-phase durations (ns), median / max over blocks:")
for k in range(1, 7):
    dd = a[:, k] - a[:, k - 1]
    print(f"  phase {k - 1}->{k}: median {np.median(dd):8.0f}  p90 {np.percentile(dd, 90):8.0f}  max {dd.max():8.0f}")
print("block start: median", np.median(a[:, 0]), "max", a[:, 0].max(), "; block end: median", np.median(a[:, 6]), "max", a[:, 6].max())
order = np.argsort(a[:, 0])
for b in list(order[:3]) + list(order[-3:]):
    print(b, a[b])
if len(sys.argv) > 2:
    b = out.reshape(-1, 8)[:nblk].astype(np.int64)
    b[:, :7] -= t0
    tk = b[:, 7]
    o = np.argsort(tk)
    pub = b[o, 2]; seen = b[o, 5]; done = b[o, 3]
    half = nblk // 2
    for name, sl in (("seg0", slice(0, half)), ("seg1", slice(half, nblk))):
        p, sn, dn = pub[sl], seen[sl], done[sl]
        pm = np.maximum.accumulate(p)
        lag = sn[1:] - pm[:-1]
        print(name, "publish median", np.median(p), "max", p.max(), "| seen - max(publish of predecessors): median", np.median(lag), "p90", np.percentile(lag, 90), "max", lag.max(),
              "| done - seen median", np.median(dn - sn))
        print("   tiles:", [(int(i), int(p[i]), int(sn[i]), int(dn[i])) for i in (0, 1, 2, 40, 80, 120, 170)])
