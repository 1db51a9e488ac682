"""Per-kernel breakdown of one BASELINE config-4 bundle (15 keyframes x 100k points, production keyframe settings)."""
import json
import sys
import time

sys.path.insert(0, ".")
from dmsa_lidar_slam_b200 import DmsaOptimSettings, MapManagement, synth  # noqa: E402
from dmsa_lidar_slam_b200.distributed import KEYFRAME_SETTINGS  # noqa: E402

split = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sm = synth.make_keyframe_submap(n_keyframes=15, n_points=100000, seed=4)
s = DmsaOptimSettings(**dict(KEYFRAME_SETTINGS, gauss_split=split))
kf = MapManagement.from_submap(sm)
p0 = kf.getPoses()
for _ in range(3):
    kf.setRelativePoses(p0["rel_orient"], p0["rel_transl"])
    d = kf.iteration(s)
kf.profileEnable(True)
t0 = time.perf_counter()
n = 5
for _ in range(n):
    kf.setRelativePoses(p0["rel_orient"], p0["rel_transl"])
    d = kf.iteration(s)
kf.ctx.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
prof = kf.profileRead()
print(json.dumps(dict(gauss_split=split, G=d["num_gaussians"], wall_ms=wall, breakdown={k: round(v[0] / n, 4) for k, v in prof.items() if v[1]})))
