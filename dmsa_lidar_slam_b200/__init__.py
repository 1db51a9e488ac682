"""dmsa_lidar_slam_b200 — B200-native (sm_100a) DMSA inner loop behind the reference's optimizer API.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C-ABI of include/dmsa_b200.h),
`api.py` (host-side mirror of DmsaOptimizer / OptimizablePointSet over the C-ABI), `synth.py`
(deterministic synthetic windows of the BASELINE shapes) and `build.py` (in-tree nvcc build).
"""
from .api import (ContinuousTrajectory, DmsaError, DmsaOptimizer, DmsaOptimSettings, MapManagement,  # noqa: F401
                  OptimizablePointSet, PreProcessor, PreprocessConfig, decode_pointcloud2, format_tum_pose, load_library,
                  pc2_layout_for_sensor, rand_sequence, save_pcd_ascii)
