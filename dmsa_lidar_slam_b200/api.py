"""Host-side mirror of the reference's optimizer interface over the C-ABI (include/dmsa_b200.h).

Names follow the reference so that parity tests read like the reference's own call sites:

    DmsaOptimSettings            <-> struct DmsaOptimSettings            DmsaOptimizer.h:25-39
    DmsaOptimizer.optimizeSet    <-> DmsaOptimizer<PointT>::optimizeSet  DmsaOptimizer.h:54-150
    ContinuousTrajectory         <-> hot members of ContinuousTrajectory ContinuousTrajectory.h:24-346
    MapManagement                <-> hot members of MapManagement        MapManagement.h:20-252

The shared library is hand-written CUDA for sm_100a; there is NO CPU fallback.  Importing this module without the
built library, or creating a context without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .synth import POINT_NORMAL, POINT_STAMP_ID

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMSA_B200_LIB") or os.path.join(_HERE, "lib", "libdmsa_b200.so")  # (override: experiment builds only)

STOP_REASONS = {0: "max_iter", 1: "few_gaussians", 2: "nan", 3: "no_improvement", 4: "epsilon"}


class DmsaError(RuntimeError):
    pass


class DmsaOptimSettings(C.Structure):
    """Field-for-field mirror of DmsaOptimSettings (DmsaOptimizer.h:25-39); defaults are the reference's."""

    _fields_ = [
        ("num_iter", C.c_int32), ("epsilon", C.c_double), ("use_analytic_jacobi", C.c_int32), ("step_length_optim", C.c_double),
        ("max_step", C.c_double), ("gauss_split", C.c_int32), ("grid_size_1_factor", C.c_float), ("grid_size_2_factor", C.c_float),
        ("min_num_points_per_set", C.c_int32), ("min_num_gaussians", C.c_int32), ("lambda_diag", C.c_float), ("use_centralization", C.c_int32),
    ]

    def __init__(self, **kw):
        super().__init__(15, 1e-5, 0, 0.05, 0.01, 0, 2.0, 5.0, 6, 30, 0.00001, 1)
        for k, v in kw.items():
            if k not in dict(self._fields_):
                raise KeyError(k)
            setattr(self, k, v)


class Report(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("stop_reason", C.c_int32), ("num_gaussians", C.c_int32), ("num_memberships", C.c_int64),
        ("num_extra", C.c_int32), ("best_step", C.c_int32), ("error0", C.c_double), ("step_norm", C.c_double),
    ]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["stop"] = STOP_REASONS.get(self.stop_reason, "?")
        return d


class Pc2Layout(C.Structure):
    """Where x / y / z / time / ring sit in one sensor_msgs/PointCloud2 point record (src/dmsa_slam_ros.cpp:399-483)."""

    _fields_ = [("point_step", C.c_int32), ("x_offset", C.c_int32), ("y_offset", C.c_int32), ("z_offset", C.c_int32),
                ("stamp_offset", C.c_int32), ("stamp_type", C.c_int32), ("ring_offset", C.c_int32), ("ring_type", C.c_int32)]


class PreprocessConfig(C.Structure):
    """The fields of Config (Config.h:24-58) that preProcess reads (DmsaSlam.h:570-634); defaults are the reference's."""

    _fields_ = [("max_num_points_per_scan", C.c_int32), ("min_dist_ds", C.c_float), ("min_dist", C.c_float), ("lidar_to_imu", C.c_float * 16)]

    def __init__(self, max_num_points_per_scan=3000, minDistDS=30.0, min_dist=0.0, lidarToImuTform=None):
        T = np.eye(4, dtype=np.float32) if lidarToImuTform is None else np.asarray(lidarToImuTform, dtype=np.float32).reshape(4, 4)
        super().__init__(int(max_num_points_per_scan), float(minDistDS), float(min_dist), (C.c_float * 16)(*T.T.ravel()))  # column-major


_lib = None


def load_library():
    """Loads libdmsa_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DmsaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    P = C.POINTER
    sig = {
        "dmsa_b200_create": (i32, [P(vp), i32, vp]),
        "dmsa_b200_destroy": (None, [vp]),
        "dmsa_b200_last_error": (C.c_char_p, [vp]),
        "dmsa_b200_version": (i32, []),
        "dmsa_b200_fuse_threshold": (i32, []),
        "dmsa_b200_launch_count": (i64, [vp]),
        "dmsa_b200_synchronize": (i32, [vp]),
        "dmsa_b200_traj_init": (i32, [vp, f64, f64, i32, i32, f64]),
        "dmsa_b200_traj_init_window": (i32, [vp, f64, f64, i32, i32, f64]),
        "dmsa_b200_spd_solve_dev": (i32, [vp, P(DmsaOptimSettings), vp, i32, vp, vp]),
        "dmsa_b200_spd_solve": (i32, [vp, P(DmsaOptimSettings), vp, i32, vp, P(i32)]),
        "dmsa_b200_bundle_jacobian": (i32, [vp, P(DmsaOptimSettings), vp, i32, vp, i32]),
        "dmsa_b200_bundle_line_search": (i32, [vp, vp, vp, vp]),
        "dmsa_b200_bundle_verify": (i32, [vp, P(i32), P(i32)]),
        "dmsa_b200_all_reduce": (i32, [vp, vp, i64]),
        "dmsa_b200_comm_unique_id": (i32, [vp]),
        "dmsa_b200_comm_init": (i32, [vp, vp, i32, i32]),
        "dmsa_b200_comm_destroy": (i32, [vp]),
        "dmsa_b200_collective_count": (i64, [vp]),
        "dmsa_b200_traj_get_dense_poses": (i32, [vp, vp, vp]),
        "dmsa_b200_traj_register_scans": (i32, [vp, i32, P(vp), P(i64), P(f32)]),
        "dmsa_b200_traj_add_static_points": (i32, [vp, vp, i64]),
        "dmsa_b200_traj_remove_static_points": (i32, [vp]),
        "dmsa_b200_traj_get_timing": (i32, [vp, P(i32), P(f64), vp, vp, vp]),
        "dmsa_b200_traj_get_tform_ids": (i32, [vp, vp]),
        "dmsa_b200_traj_set_imu_factors": (i32, [vp, vp, vp, vp, vp, f64, vp]),
        "dmsa_b200_kf_init": (i32, [vp, i32]),
        "dmsa_b200_kf_set_keyframe": (i32, [vp, i32, vp, vp, i64, f32]),
        "dmsa_b200_kf_commit": (i32, [vp]),
        "dmsa_b200_kf_set_gravity_terms": (i32, [vp, vp, vp, f64]),
        "dmsa_b200_kf_set_odometry_terms": (i32, [vp, vp, vp, f64]),
        "dmsa_b200_set_relative_poses": (i32, [vp, vp, vp]),
        "dmsa_b200_get_poses": (i32, [vp, vp, vp, vp, vp]),
        "dmsa_b200_num_params": (i32, [vp]),
        "dmsa_b200_get_pose_parameters": (i32, [vp, vp]),
        "dmsa_b200_set_pose_parameters": (i32, [vp, vp]),
        "dmsa_b200_centralize": (i32, [vp]),
        "dmsa_b200_decentralize": (i32, [vp]),
        "dmsa_b200_update_global_points": (i32, [vp]),
        "dmsa_b200_num_points": (i64, [vp]),
        "dmsa_b200_get_global_points": (i32, [vp, vp, vp]),
        "dmsa_b200_traj_get_dense_tforms": (i32, [vp, vp]),
        "dmsa_b200_build_sets": (i32, [vp, P(DmsaOptimSettings), P(i32), P(i64)]),
        "dmsa_b200_get_sets": (i32, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "dmsa_b200_get_voxel_keys": (i32, [vp, i32, vp, vp, P(i32)]),
        "dmsa_b200_eval_cost": (i32, [vp, vp, i32, vp]),
        "dmsa_b200_cost_jacobian": (i32, [vp, vp, vp, P(f64), vp, vp]),
        "dmsa_b200_iteration": (i32, [vp, P(DmsaOptimSettings), P(i32), P(Report), vp, vp]),
        "dmsa_b200_optimize": (i32, [vp, P(DmsaOptimSettings), P(Report)]),
        "dmsa_b200_set_mean_mode": (i32, [vp, i32]),
        "dmsa_b200_profile_enable": (i32, [vp, i32]),
        "dmsa_b200_profile_num": (i32, []),
        "dmsa_b200_profile_name": (C.c_char_p, [i32]),
        "dmsa_b200_profile_read": (i32, [vp, i32, P(f64), P(i64)]),
        "dmsa_b200_set_shard": (i32, [vp, i32, i32]),
        "dmsa_b200_cost_jacobian_dev": (i32, [vp, vp]),
        "dmsa_b200_line_search_costs_dev": (i32, [vp, vp, vp]),
        "dmsa_b200_lm_solve": (i32, [P(DmsaOptimSettings), vp, i32, i32, vp, P(i32)]),
        "dmsa_b200_set_lm_solver": (i32, [vp, i32]),
        "dmsa_b200_set_pair_mode": (i32, [vp, i32]),
        "dmsa_b200_get_batch_tables": (i32, [vp, vp, vp]),
        "dmsa_b200_set_run_ahead": (i32, [vp, i32]),
        "dmsa_b200_rand_sequence": (i32, [C.c_uint32, i64, vp]),
        "dmsa_b200_grid_downsample": (i32, [vp, vp, i64, i32, C.c_float, C.c_uint32, vp, P(i64)]),
        "dmsa_b200_downsample_global_points": (i32, [vp, C.c_float, C.c_uint32, vp, P(i64)]),
        "dmsa_b200_preprocess_scan": (i32, [vp, vp, i64, P(PreprocessConfig), C.c_uint32, vp, P(i64), P(C.c_float)]),
        "dmsa_b200_estimate_normals": (i32, [vp, vp, i64, vp, C.c_float, vp]),
        "dmsa_b200_pc2_layout_for_sensor": (i32, [C.c_char_p, vp, i32, i32, P(Pc2Layout)]),
        "dmsa_b200_decode_pointcloud2": (i32, [vp, vp, i64, P(Pc2Layout), f64, f64, vp]),
        "dmsa_b200_relative2global": (i32, [i32, vp, vp, vp, vp]),
        "dmsa_b200_format_tum_pose": (i32, [f64, vp, vp, C.c_char_p, i32]),
        "dmsa_b200_save_pcd_ascii": (i32, [C.c_char_p, vp, i64]),
        "dmsa_b200_select_static_points": (i32, [vp, vp, i64, vp, C.c_float, vp, P(i64)]),
        "dmsa_b200_overlap": (i32, [vp, vp, i64, C.c_float, P(C.c_float)]),
        "dmsa_b200_lm_solve_device": (i32, [vp, P(DmsaOptimSettings), vp, i32, vp, P(i32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "dmsa_b200_create", "dmsa_b200_destroy", "dmsa_b200_last_error", "dmsa_b200_version", "dmsa_b200_fuse_threshold", "dmsa_b200_launch_count", "dmsa_b200_synchronize",
    "dmsa_b200_traj_init", "dmsa_b200_traj_register_scans", "dmsa_b200_traj_add_static_points", "dmsa_b200_traj_remove_static_points",
    "dmsa_b200_traj_get_timing", "dmsa_b200_traj_get_tform_ids", "dmsa_b200_traj_set_imu_factors", "dmsa_b200_kf_init",
    "dmsa_b200_kf_set_keyframe", "dmsa_b200_kf_commit", "dmsa_b200_kf_set_gravity_terms", "dmsa_b200_kf_set_odometry_terms",
    "dmsa_b200_set_relative_poses", "dmsa_b200_get_poses", "dmsa_b200_num_params", "dmsa_b200_get_pose_parameters",
    "dmsa_b200_set_pose_parameters", "dmsa_b200_centralize", "dmsa_b200_decentralize", "dmsa_b200_update_global_points",
    "dmsa_b200_num_points", "dmsa_b200_get_global_points", "dmsa_b200_traj_get_dense_tforms", "dmsa_b200_build_sets", "dmsa_b200_get_sets",
    "dmsa_b200_get_voxel_keys", "dmsa_b200_eval_cost", "dmsa_b200_cost_jacobian", "dmsa_b200_iteration", "dmsa_b200_optimize",
    "dmsa_b200_set_mean_mode", "dmsa_b200_profile_enable", "dmsa_b200_profile_num", "dmsa_b200_profile_name", "dmsa_b200_profile_read",
    "dmsa_b200_set_shard", "dmsa_b200_cost_jacobian_dev", "dmsa_b200_line_search_costs_dev", "dmsa_b200_lm_solve",
    "dmsa_b200_set_lm_solver", "dmsa_b200_lm_solve_device", "dmsa_b200_set_pair_mode", "dmsa_b200_get_batch_tables",
    "dmsa_b200_relative2global", "dmsa_b200_pc2_layout_for_sensor", "dmsa_b200_decode_pointcloud2", "dmsa_b200_format_tum_pose", "dmsa_b200_save_pcd_ascii",
    "dmsa_b200_set_run_ahead", "dmsa_b200_rand_sequence", "dmsa_b200_grid_downsample", "dmsa_b200_downsample_global_points", "dmsa_b200_preprocess_scan", "dmsa_b200_estimate_normals",
    "dmsa_b200_select_static_points", "dmsa_b200_overlap", "dmsa_b200_traj_init_window", "dmsa_b200_traj_get_dense_poses",
    "dmsa_b200_spd_solve_dev", "dmsa_b200_spd_solve", "dmsa_b200_bundle_jacobian", "dmsa_b200_bundle_line_search", "dmsa_b200_bundle_verify",
    "dmsa_b200_all_reduce", "dmsa_b200_comm_unique_id", "dmsa_b200_comm_init", "dmsa_b200_comm_destroy", "dmsa_b200_collective_count",
]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Context:
    """One C-ABI context (== one optimizer instance + one staged point-set model)."""

    def __init__(self, device=0, stream=None):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.dmsa_b200_create(C.byref(h), int(device), C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise DmsaError({3: "no CUDA device: dmsa_b200 has no CPU fallback"}.get(rc, f"dmsa_b200_create failed ({rc})"))
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.dmsa_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise DmsaError(f"dmsa_b200 error {rc}: {self.L.dmsa_b200_last_error(self.h).decode()}")

    @property
    def launch_count(self):
        return self.L.dmsa_b200_launch_count(self.h)

    def synchronize(self):
        self._ck(self.L.dmsa_b200_synchronize(self.h))


class OptimizablePointSet:
    """Common part of the two models (OptimizablePointSet.h:18-56): pose parameters, globalPoints, minGridSize."""

    def __init__(self, device=0, stream=None):
        self.ctx = _Context(device, stream)
        self.L = self.ctx.L
        self.h = self.ctx.h
        self.n_poses = 0

    # ---- getPoseParameters / setPoseParameters
    @property
    def numParams(self):
        return self.L.dmsa_b200_num_params(self.h)

    def getPoseParameters(self):
        p = np.zeros(self.numParams)
        self.ctx._ck(self.L.dmsa_b200_get_pose_parameters(self.h, _p(p)))
        return p

    def setPoseParameters(self, params):
        p = _c64(params)
        assert p.size == self.numParams
        self.ctx._ck(self.L.dmsa_b200_set_pose_parameters(self.h, _p(p)))

    def setRelativePoses(self, rel_orient, rel_transl):
        """rel_orient / rel_transl: 3 x n_poses (Poses.h:19-20)."""
        ro = np.ascontiguousarray(np.asarray(rel_orient, dtype=np.float64).T)
        rt = np.ascontiguousarray(np.asarray(rel_transl, dtype=np.float64).T)
        assert ro.shape == (self.n_poses, 3) and rt.shape == (self.n_poses, 3)
        self.ctx._ck(self.L.dmsa_b200_set_relative_poses(self.h, _p(ro), _p(rt)))

    def getPoses(self):
        n = self.n_poses
        out = [np.zeros((n, 3)) for _ in range(4)]
        self.ctx._ck(self.L.dmsa_b200_get_poses(self.h, *[_p(o) for o in out]))
        return dict(rel_orient=out[0].T.copy(), rel_transl=out[1].T.copy(), glob_orient=out[2].T.copy(), glob_transl=out[3].T.copy())

    def centralize(self):
        self.ctx._ck(self.L.dmsa_b200_centralize(self.h))

    def decentralize(self):
        self.ctx._ck(self.L.dmsa_b200_decentralize(self.h))

    # ---- globalPoints
    def updateGlobalPoints(self):
        self.ctx._ck(self.L.dmsa_b200_update_global_points(self.h))

    @property
    def numPoints(self):
        return self.L.dmsa_b200_num_points(self.h)

    def globalPoints(self, normals=False):
        N = self.numPoints
        xyzw = np.zeros((N, 4), dtype=np.float32)
        nrm = np.zeros((N, 4), dtype=np.float32) if normals else None
        self.ctx._ck(self.L.dmsa_b200_get_global_points(self.h, _p(xyzw), _p(nrm)))
        return (xyzw, nrm) if normals else xyzw

    # ---- step-by-step hot path (debug / parity surface)
    def buildSets(self, settings):
        G, M = C.c_int32(), C.c_int64()
        self.ctx._ck(self.L.dmsa_b200_build_sets(self.h, C.byref(settings), C.byref(G), C.byref(M)))
        self._G, self._M = G.value, M.value
        return G.value, M.value

    def getSets(self):
        G, M = self._G, self._M
        offs = np.zeros(G + 1, dtype=np.int64)
        members = np.zeros(M, dtype=np.int32)
        info = np.zeros((G, 9), dtype=np.float32)
        w = np.zeros(G, dtype=np.float32)
        level = np.zeros(G, dtype=np.int32)
        key = np.zeros((G, 3), dtype=np.int32)
        sub = np.zeros(G, dtype=np.int32)
        self.ctx._ck(self.L.dmsa_b200_get_sets(self.h, _p(offs), _p(members), _p(info), _p(w), _p(level), _p(key), _p(sub)))
        return dict(G=G, M=M, offs=offs, members=members, info=info, w=w, level=level, key=key, sub=sub)

    def voxelKeys(self, level):
        N = self.numPoints
        keys = np.zeros((N, 3), dtype=np.int32)
        lo = np.zeros(3, dtype=np.int64)
        depth = C.c_int32()
        self.ctx._ck(self.L.dmsa_b200_get_voxel_keys(self.h, int(level), _p(keys), _p(lo), C.byref(depth)))
        return keys, lo, depth.value

    def numExtra(self):
        return 0

    def evalCost(self, params):
        """params: V x P -> e: V x (G+E)  (V cost evaluations, DmsaOptimizer.h:234-273)."""
        p = _c64(np.atleast_2d(params))
        V = p.shape[0]
        R = self._G + self.numExtra()
        e = np.zeros((V, R))
        self.ctx._ck(self.L.dmsa_b200_eval_cost(self.h, _p(p), V, _p(e)))
        return e

    def costJacobian(self, with_rows=False):
        P = self.numParams
        R = self._G + self.numExtra()
        H = np.zeros((P, P))
        g = np.zeros(P)
        err0 = C.c_double()
        e0 = np.zeros(R) if with_rows else None
        J = np.zeros((P, R)) if with_rows else None
        self.ctx._ck(self.L.dmsa_b200_cost_jacobian(self.h, _p(H), _p(g), C.byref(err0), _p(e0), _p(J)))
        out = dict(H=H, g=g, err0=err0.value)
        if with_rows:
            out["e0"] = e0
            out["J"] = J.T.copy()
        return out

    def iteration(self, settings):
        P = self.numParams
        stop = C.c_int32()
        rep = Report()
        step = np.zeros(P)
        ls = np.zeros(9)
        self.ctx._ck(self.L.dmsa_b200_iteration(self.h, C.byref(settings), C.byref(stop), C.byref(rep), _p(step), _p(ls)))
        d = rep.asdict()
        d.update(step=step, ls_cost=ls, stop_reason=stop.value, stop=STOP_REASONS[stop.value])
        self._G = rep.num_gaussians
        return d

    def setMeanMode(self, mode):
        """0: order-free exactly-rounded per-set mean (default, fast); 1: the reference's sequential float accumulation."""
        self.ctx._ck(self.L.dmsa_b200_set_mean_mode(self.h, int(mode)))

    def selectStaticPoints(self, cloud, pos, max_dist):
        """DmsaSlam.h:304-339 for one keyframe cloud (48-byte PointNormal records, world frame): (selected uint8[n], currOverlap)."""
        c = np.ascontiguousarray(cloud)
        assert c.dtype.itemsize == 48, "pcl::PointNormal records (48 bytes)"
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        sel = np.zeros(len(c), dtype=np.uint8)
        cnt = C.c_int64(0)
        self.ctx._ck(self.L.dmsa_b200_select_static_points(self.h, _p(c), len(c), _p(pos), float(max_dist), _p(sel), C.byref(cnt)))
        return sel, int(cnt.value)

    def overlap(self, pc1_xyzw, max_dist):
        """getOverlap(pc1, window cloud, max_dist), DmsaSlam.h:377-414."""
        a = np.ascontiguousarray(pc1_xyzw, dtype=np.float32).reshape(-1, 4)
        out = C.c_float(0.0)
        self.ctx._ck(self.L.dmsa_b200_overlap(self.h, _p(a), len(a), float(max_dist), C.byref(out)))
        return float(out.value)

    def batchTables(self):
        """The float transform table of the batch evaluated last: (rows + 1, V, 12)."""
        dims = np.zeros(2, dtype=np.int32)
        self.ctx._ck(self.L.dmsa_b200_get_batch_tables(self.h, None, _p(dims)))
        out = np.zeros((int(dims[0]), int(dims[1]), 12), dtype=np.float32)
        self.ctx._ck(self.L.dmsa_b200_get_batch_tables(self.h, _p(out), _p(dims)))
        return out

    def setRunAhead(self, on):
        """optimizeSet: run-ahead loop (default) or one loop body at a time; bit-identical."""
        self.ctx._ck(self.L.dmsa_b200_set_run_ahead(self.h, int(bool(on))))

    def setPairMode(self, mode):
        """1: pair-packed FP32x2 cost kernels for the forward-difference batch (default); 0: scalar kernels (bit-identical)."""
        self.ctx._ck(self.L.dmsa_b200_set_pair_mode(self.h, int(mode)))

    def setLmSolver(self, mode):
        """0: LM step on the device for P <= 128 (default; no host round trip inside a loop body); 1: host solver (bit-identical);
        2: device Cholesky for every P <= 1024 (agrees to the conditioning of the system, not bit for bit)."""
        self.ctx._ck(self.L.dmsa_b200_set_lm_solver(self.h, int(mode)))

    def lmSolveDevice(self, settings, hg, n_params):
        """Device twin of api.lm_solve(.., explicit_inverse=1): (step, has_nan) for a host [H | g | err0] buffer."""
        hg = np.ascontiguousarray(hg, dtype=np.float64)
        step = np.zeros(int(n_params))
        nan = C.c_int32(0)
        self.ctx._ck(self.L.dmsa_b200_lm_solve_device(self.h, C.byref(settings), _p(hg), int(n_params), _p(step), C.byref(nan)))
        return step, int(nan.value)

    def spdSolve(self, settings, hg, n_params):
        """Device Cholesky LM step of the bundle extension on a host [H | g | err0] buffer: (step, flag 0 ok / 1 NaN / 2 not SPD)."""
        hg = np.ascontiguousarray(hg, dtype=np.float64)
        step = np.zeros(int(n_params))
        flag = C.c_int32(0)
        self.ctx._ck(self.L.dmsa_b200_spd_solve(self.h, C.byref(settings), _p(hg), int(n_params), _p(step), C.byref(flag)))
        return step, int(flag.value)

    def spdSolveDev(self, settings, hg_dev_ptr, n_params, step_dev_ptr, tail_dev_ptr):
        self.ctx._ck(self.L.dmsa_b200_spd_solve_dev(self.h, C.byref(settings), C.c_void_p(hg_dev_ptr), int(n_params), C.c_void_p(step_dev_ptr),
                                                    C.c_void_p(tail_dev_ptr)))

    def bundleJacobian(self, settings, idx_dev_ptr, P_global, ghg_dev_ptr, sync_build=False):
        self.ctx._ck(self.L.dmsa_b200_bundle_jacobian(self.h, C.byref(settings), C.c_void_p(idx_dev_ptr), int(P_global), C.c_void_p(ghg_dev_ptr),
                                                      int(bool(sync_build))))

    def bundleLineSearch(self, gstep_dev_ptr, idx_dev_ptr, gls_dev_ptr):
        self.ctx._ck(self.L.dmsa_b200_bundle_line_search(self.h, C.c_void_p(gstep_dev_ptr), C.c_void_p(idx_dev_ptr), C.c_void_p(gls_dev_ptr)))

    def bundleVerify(self):
        G, redo = C.c_int32(0), C.c_int32(0)
        self.ctx._ck(self.L.dmsa_b200_bundle_verify(self.h, C.byref(G), C.byref(redo)))
        self._G = G.value
        return G.value, bool(redo.value)

    def allReduce(self, dev_ptr, count):
        self.ctx._ck(self.L.dmsa_b200_all_reduce(self.h, C.c_void_p(dev_ptr), int(count)))

    def profileEnable(self, on=True):
        self.ctx._ck(self.L.dmsa_b200_profile_enable(self.h, int(bool(on))))

    def profileRead(self):
        """{kernel-or-phase name: (total device ms, launches)} measured with CUDA events on the launching stream."""
        out = {}
        for i in range(self.L.dmsa_b200_profile_num()):
            ms, cnt = C.c_double(), C.c_int64()
            self.ctx._ck(self.L.dmsa_b200_profile_read(self.h, i, C.byref(ms), C.byref(cnt)))
            out[self.L.dmsa_b200_profile_name(i).decode()] = (ms.value, cnt.value)
        return out

    def costJacobianDev(self, hg_dev_ptr):
        """Partial [H | g | err0] of this context's sets into caller-owned DEVICE memory (P*P + P + 1 doubles)."""
        self.ctx._ck(self.L.dmsa_b200_cost_jacobian_dev(self.h, C.c_void_p(hg_dev_ptr)))

    def lineSearchCostsDev(self, step, ls_dev_ptr):
        """Partial costs of the 9 trial points p + 0.1 k step into caller-owned DEVICE memory (9 doubles)."""
        st = _c64(step)
        self.ctx._ck(self.L.dmsa_b200_line_search_costs_dev(self.h, _p(st), C.c_void_p(ls_dev_ptr)))

    def setShard(self, rank, world):
        self.ctx._ck(self.L.dmsa_b200_set_shard(self.h, int(rank), int(world)))

    def commInit(self, unique_id: bytes, rank, world):
        """Row sharding across ranks with the in-library NCCL exchange (every rank stages the same set)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.ctx._ck(self.L.dmsa_b200_comm_init(self.h, buf, int(rank), int(world)))

    def commDestroy(self):
        self.ctx._ck(self.L.dmsa_b200_comm_destroy(self.h))

    @property
    def collective_count(self):
        return self.L.dmsa_b200_collective_count(self.h)


class ContinuousTrajectory(OptimizablePointSet):
    """Sliding-window model.  Mirrors ContinuousTrajectory: initTraj, registerPcBuffer, add/removeStaticPoints."""

    def initTraj(self, t_min, t_max, numControlPoses, useImu=False, dtResIn=1e-3):
        self.n_poses = int(numControlPoses)
        self.ctx._ck(self.L.dmsa_b200_traj_init(self.h, float(t_min), float(t_max), self.n_poses, int(bool(useImu)), float(dtResIn)))
        self._use_imu = bool(useImu)
        self._imu_set = False

    def registerPcBuffer(self, scans, grid_sizes):
        """scans: list of POINT_STAMP_ID arrays in chronological order (RingBuffer.h:31-37)."""
        n = len(scans)
        scans = [np.ascontiguousarray(s, dtype=POINT_STAMP_ID) for s in scans]
        self.ctx._keep = scans
        ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in scans])
        sizes = (C.c_int64 * n)(*[len(s) for s in scans])
        gs = (C.c_float * n)(*[float(g) for g in grid_sizes])
        self.ctx._ck(self.L.dmsa_b200_traj_register_scans(self.h, n, ptrs, sizes, gs))

    def addStaticPoints(self, pts):
        pts = np.ascontiguousarray(pts, dtype=POINT_STAMP_ID)
        self.ctx._ck(self.L.dmsa_b200_traj_add_static_points(self.h, _p(pts), len(pts)))

    def removeStaticPoints(self):
        self.ctx._ck(self.L.dmsa_b200_traj_remove_static_points(self.h))

    def timing(self):
        nt = C.c_int32()
        hor = C.c_double()
        self.ctx._ck(self.L.dmsa_b200_traj_get_timing(self.h, C.byref(nt), C.byref(hor), None, None, None))
        stamps = np.zeros(self.n_poses)
        tt = np.zeros(nt.value)
        pi = np.zeros(self.n_poses, dtype=np.int32)
        self.ctx._ck(self.L.dmsa_b200_traj_get_timing(self.h, C.byref(nt), C.byref(hor), _p(stamps), _p(tt), _p(pi)))
        return dict(n_total=nt.value, horizon=hor.value, stamps=stamps, traj_time=tt, param_indices=pi)

    def tformIdPerPoint(self, n_scan_points):
        out = np.zeros(n_scan_points, dtype=np.int32)
        self.ctx._ck(self.L.dmsa_b200_traj_get_tform_ids(self.h, _p(out)))
        return out

    def denseTforms(self):
        nt = self.timing()["n_total"]
        out = np.zeros((nt, 12), dtype=np.float32)
        self.ctx._ck(self.L.dmsa_b200_traj_get_dense_tforms(self.h, _p(out)))
        return out

    def denseGlobalPoses(self):
        """denseGlobalPoses (ContinuousTrajectory.h:29): (orientations 3 x n_total, translations 3 x n_total) doubles."""
        nt = self.timing()["n_total"]
        o, t = np.zeros((nt, 3)), np.zeros((nt, 3))
        self.ctx._ck(self.L.dmsa_b200_traj_get_dense_poses(self.h, _p(o), _p(t)))
        return o.T.copy(), t.T.copy()

    def setImuFactors(self, preint_rot, preint_pos, preint_vel, cov_inv, balancing=0.001, gravity=(0.0, 0.0, -9.805)):
        a = [_c64(preint_rot), _c64(preint_pos), _c64(preint_vel), _c64(cov_inv), _c64(gravity)]
        self.ctx._ck(self.L.dmsa_b200_traj_set_imu_factors(self.h, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), float(balancing), _p(a[4])))
        self._imu_set = True

    def numExtra(self):
        return (self.n_poses - 1) if (self._use_imu and self._imu_set) else 0

    @classmethod
    def from_window(cls, win, device=0, stream=None, use_imu=False):
        """Stages a synthetic window (synth.make_sliding_window) exactly like DmsaSlam::prepareTrajectoryForOptimization would."""
        t = cls(device, stream)
        t.initTraj(win["t_min"], win["t_max"], win["n_poses"], use_imu, win["dt_res"])
        t.registerPcBuffer(win["scans"], win["grid_sizes"])
        if len(win["static"]):
            t.addStaticPoints(win["static"])
        t.setRelativePoses(win["rel_orient"], win["rel_transl"])
        return t


class MapManagement(OptimizablePointSet):
    """Keyframe-submap model.  Mirrors the hot members of MapManagement (keyframe clouds + poses + factors)."""

    def __init__(self, n_max=30, device=0, stream=None):
        super().__init__(device, stream)
        self.n_poses = int(n_max)
        self.ctx._ck(self.L.dmsa_b200_kf_init(self.h, self.n_poses))
        self._grav = self._odom = False

    def setKeyframe(self, k, cloud, ring_ids, grid_size):
        cloud = np.ascontiguousarray(cloud, dtype=POINT_NORMAL)
        ring = np.ascontiguousarray(ring_ids, dtype=np.int32)
        self.ctx._ck(self.L.dmsa_b200_kf_set_keyframe(self.h, int(k), _p(cloud), _p(ring), len(cloud), float(grid_size)))

    def commit(self):
        self.ctx._ck(self.L.dmsa_b200_kf_commit(self.h))

    def setGravityTerms(self, measured_gravity, plausible, balance=1.0):
        a = _c64(measured_gravity)
        b = np.ascontiguousarray(plausible, dtype=np.int32)
        self.ctx._ck(self.L.dmsa_b200_kf_set_gravity_terms(self.h, _p(a), _p(b), float(balance)))
        self._grav = True

    def setOdometryTerms(self, rel_transl, rel_orient_mat, balance=1000.0):
        a, b = _c64(rel_transl), _c64(rel_orient_mat)
        self.ctx._ck(self.L.dmsa_b200_kf_set_odometry_terms(self.h, _p(a), _p(b), float(balance)))
        self._odom = True

    def numExtra(self):
        return (self.n_poses if self._grav else 0) + (self.n_poses - 1 if self._odom else 0)

    @classmethod
    def from_submap(cls, sm, device=0, stream=None):
        m = cls(sm["n_keyframes"], device, stream)
        for k, (c, r, g) in enumerate(zip(sm["clouds"], sm["rings"], sm["grid_sizes"])):
            m.setKeyframe(k, c, r, g)
        m.commit()
        m.setRelativePoses(sm["rel_orient"], sm["rel_transl"])
        return m


class DmsaOptimizer:
    """DmsaOptimizer<PointT> (DmsaOptimizer.h:41-182) — optimizeSet mutates the set in place (poses, globalPoints)."""

    def optimizeSet(self, pointSetToOptimize: OptimizablePointSet, settings: DmsaOptimSettings | None = None):
        settings = settings or DmsaOptimSettings()
        rep = Report()
        s = pointSetToOptimize
        s.ctx._ck(s.L.dmsa_b200_optimize(s.h, C.byref(settings), C.byref(rep)))
        self.last_report = rep.asdict()
        s._G = rep.num_gaussians
        return self.last_report


def comm_unique_id() -> bytes:
    """128-byte NCCL id for dmsa_b200_comm_init: create on rank 0, broadcast to the other ranks."""
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.dmsa_b200_comm_unique_id(buf)
    if rc != 0:
        raise DmsaError(f"dmsa_b200_comm_unique_id failed ({rc}): is libnccl.so.2 loadable?")
    return buf.raw


def lm_solve(settings, hg, n_params, explicit_inverse=True):
    """step = -alpha (H + lambda I)^-1 g and clamp (DmsaOptimizer.h:107-128); explicit_inverse=True is the reference's arithmetic."""
    L = load_library()
    hg = _c64(hg)
    step = np.zeros(n_params)
    nan = C.c_int32()
    rc = L.dmsa_b200_lm_solve(C.byref(settings), _p(hg), int(n_params), int(explicit_inverse), _p(step), C.byref(nan))
    if rc != 0:
        raise DmsaError(f"dmsa_b200_lm_solve failed ({rc})")
    return step, bool(nan.value)


# ---- pre-processing and normal estimation (helpers.h:67-182, DmsaSlam.h:557-634) ----------------------
def rand_sequence(seed, n):
    """The first n values of rand() after srand(seed) (glibc), as the library generates them."""
    L = load_library()
    out = np.zeros(int(n), dtype=np.int32)
    rc = L.dmsa_b200_rand_sequence(int(seed) & 0xFFFFFFFF, int(n), _p(out))
    if rc:
        raise DmsaError("rand_sequence failed")
    return out


class PreProcessor:
    """randomGridDownsampling / preProcess / updateNormals of DmsaSlam on the device (own context unless one is passed)."""

    def __init__(self, ctx=None, device=0, stream=None):
        self.ctx = ctx if ctx is not None else _Context(device, stream)
        self.L, self.h = self.ctx.L, self.ctx.h

    def randomGridDownsampling(self, cloud, gridSize, seed):
        """cloud: structured array (PointStampId or pcl::PointNormal layout) -> (filtered cloud, picked indices)."""
        cloud = np.ascontiguousarray(cloud)
        idx = np.zeros(len(cloud), dtype=np.int32)
        n_out = C.c_int64(0)
        self.ctx._ck(self.L.dmsa_b200_grid_downsample(self.h, _p(cloud), len(cloud), cloud.dtype.itemsize, float(gridSize), int(seed) & 0xFFFFFFFF, _p(idx), C.byref(n_out)))
        idx = idx[: n_out.value]
        return cloud[idx], idx

    def preProcess(self, rawPc, config: PreprocessConfig, seed):
        """-> (filteredPc as PointStampId array, gridSize)."""
        rawPc = np.ascontiguousarray(rawPc, dtype=POINT_STAMP_ID)
        out = np.zeros(len(rawPc), dtype=POINT_STAMP_ID)
        n_out, gs = C.c_int64(0), C.c_float(0.0)
        self.ctx._ck(self.L.dmsa_b200_preprocess_scan(self.h, _p(rawPc), len(rawPc), C.byref(config), int(seed) & 0xFFFFFFFF, _p(out), C.byref(n_out), C.byref(gs)))
        return out[: n_out.value], float(gs.value)

    def updateNormals(self, cloud, origin=(0.0, 0.0, 0.0), cell_size=0.3, with_neighbours=False):
        """cloud: pcl::PointNormal structured array; returns the cloud with normals / curvature (and the n x 6 neighbour indices)."""
        cloud = np.ascontiguousarray(cloud, dtype=POINT_NORMAL).copy()
        vp = np.asarray(origin, dtype=np.float32)
        nn = np.zeros((len(cloud), 6), dtype=np.int32) if with_neighbours else None
        self.ctx._ck(self.L.dmsa_b200_estimate_normals(self.h, _p(cloud), len(cloud), _p(vp), float(cell_size), _p(nn) if with_neighbours else None))
        return (cloud, nn) if with_neighbours else cloud


# ---- data formats (src/dmsa_slam_ros.cpp:286-291, 372-486; OutputManagement.h:80-96) -------------------
def pc2_layout_for_sensor(sensor, field_offsets, point_step):
    L = load_library()
    fo = np.ascontiguousarray(field_offsets, dtype=np.int32)
    out = Pc2Layout()
    if L.dmsa_b200_pc2_layout_for_sensor(sensor.encode(), _p(fo), len(fo), int(point_step), C.byref(out)):
        raise DmsaError(f"unknown sensor type / missing field: {sensor}")
    return out


def decode_pointcloud2(ctx, data, n_points, layout, stamp_msg, delta_t=0.0):
    """callbackPointCloud's per-point loop: msg->data (bytes / uint8 array) -> PointStampId array."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(int(n_points), dtype=POINT_STAMP_ID)
    ctx._ck(ctx.L.dmsa_b200_decode_pointcloud2(ctx.h, _p(buf), int(n_points), C.byref(layout), float(stamp_msg), float(delta_t), _p(out)))
    return out


def relative2global(rel_orient, rel_transl):
    """ConsecutivePoses.h:26-43 on 3 x n arrays (columns = poses) -> (global orientations, global translations)."""
    L = load_library()
    ro, rt = _c64(np.asarray(rel_orient).T), _c64(np.asarray(rel_transl).T)  # n x 3 row-major == 3 x n column-major
    go, gt = np.zeros_like(ro), np.zeros_like(rt)
    if L.dmsa_b200_relative2global(len(ro), _p(ro), _p(rt), _p(go), _p(gt)):
        raise DmsaError("relative2global failed")
    return np.ascontiguousarray(go.T), np.ascontiguousarray(gt.T)


def format_tum_pose(stamp, pos, orient):
    L = load_library()
    buf = C.create_string_buffer(256)
    n = L.dmsa_b200_format_tum_pose(float(stamp), _p(_c64(pos)), _p(_c64(orient)), buf, 256)
    if n < 0:
        raise DmsaError("format_tum_pose failed")
    return buf.value.decode()


def save_pcd_ascii(filename, cloud):
    c = np.ascontiguousarray(cloud, dtype=POINT_NORMAL)
    if load_library().dmsa_b200_save_pcd_ascii(str(filename).encode(), _p(c), len(c)):
        raise DmsaError(f"cannot write {filename}")
