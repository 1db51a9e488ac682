import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dmsa_lidar_slam_b200 import synth, MapManagement, DmsaOptimSettings
from dmsa_lidar_slam_b200.distributed import KeyframeBundleOptimizer

st = dict(num_iter=1, step_length_optim=0.2, max_step=0.01, min_num_points_per_set=10, min_num_gaussians=30, gauss_split=0, epsilon=1e-4)
s = DmsaOptimSettings(**st)
for (nk, npts, size, ov) in ((8, 20000, 8, 0), (8, 20000, 5, 2), (20, 20000, 8, 4), (64, 20000, 15, 8)):
    sm = synth.make_keyframe_submap(n_keyframes=nk, n_points=npts, seed=4)
    opt = KeyframeBundleOptimizer(sm, s, size, ov)
    r = opt.iteration()
    print(nk, npts, size, ov, 'bundles', opt.ranges, {k: (v if not isinstance(v, np.ndarray) else np.round(v[[0, 4, 8]], 1)) for k, v in r.items()})
    if size >= nk:
        kf = MapManagement.from_submap(sm)
        d = kf.iteration(s)
        print('   plain iteration:', d['stop'], d['error0'], d['best_step'], d['step_norm'], np.round(d['ls_cost'][[0, 4, 8]], 1))
