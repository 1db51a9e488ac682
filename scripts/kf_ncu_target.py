"""Two iterations of one BASELINE config-4 bundle (15 keyframes x 100k points, production keyframe settings): ncu target."""
import sys

sys.path.insert(0, ".")
from dmsa_lidar_slam_b200 import DmsaOptimSettings, MapManagement, synth  # noqa: E402
from dmsa_lidar_slam_b200.distributed import KEYFRAME_SETTINGS  # noqa: E402

sm = synth.make_keyframe_submap(n_keyframes=15, n_points=100000, seed=4)
s = DmsaOptimSettings(**KEYFRAME_SETTINGS)
kf = MapManagement.from_submap(sm)
p0 = kf.getPoses()
for _ in range(2):
    kf.setRelativePoses(p0["rel_orient"], p0["rel_transl"])
    d = kf.iteration(s)
print(d["num_gaussians"], d["error0"])
