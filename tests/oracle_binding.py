"""ctypes binding of the CPU oracle (oracle/libdmsa_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class OrcSettings(C.Structure):
    # DmsaOptimizer.h:25-39, same field order
    _fields_ = [
        ("num_iter", C.c_int), ("epsilon", C.c_double), ("use_analytic_jacobi", C.c_int), ("step_length_optim", C.c_double),
        ("max_step", C.c_double), ("gauss_split", C.c_int), ("grid_size_1_factor", C.c_float), ("grid_size_2_factor", C.c_float),
        ("min_num_points_per_set", C.c_int), ("min_num_gaussians", C.c_int), ("lambda_diag", C.c_float), ("use_centralization", C.c_int),
    ]


def settings(**kw):
    s = OrcSettings(15, 1e-5, 0, 0.05, 0.01, 0, 2.0, 5.0, 6, 30, 0.00001, 1)
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def build(opt="O2"):
    name = "libdmsa_oracle.so" if opt == "O2" else "libdmsa_oracle_O1.so"
    path = os.path.join(ORACLE_DIR, name)
    src = os.path.join(ORACLE_DIR, "dmsa_oracle.cpp")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, name], stdout=subprocess.DEVNULL)
    return path


_LIBS = {}


def lib(opt="O2"):
    if opt in _LIBS:
        return _LIBS[opt]
    L = C.CDLL(build(opt))
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    P = C.POINTER
    L.orc_traj_create.restype = vp
    L.orc_traj_create.argtypes = [i32, vp, i32, vp, f64, vp, vp, vp, i64, vp, vp, i64, f32, vp, vp]
    L.orc_traj_set_imu.argtypes = [vp, vp, vp, vp, vp, vp, f64, vp]
    L.orc_kf_create.restype = vp
    L.orc_kf_create.argtypes = [i32, vp, vp, vp, vp, i64, f32, vp, vp]
    L.orc_kf_set_gravity.argtypes = [vp, vp, vp, f64]
    L.orc_kf_set_odometry.argtypes = [vp, vp, vp, f64]
    L.orc_destroy.argtypes = [vp]
    L.orc_set_mode.argtypes = [vp, i32]
    L.orc_set_threads.argtypes = [vp, i32]
    L.orc_num_params.argtypes = [vp]
    L.orc_num_extra.argtypes = [vp]
    L.orc_get_params.argtypes = [vp, vp]
    L.orc_set_params.argtypes = [vp, vp]
    L.orc_get_poses.argtypes = [vp, vp, vp, vp, vp]
    L.orc_centralize.argtypes = [vp]
    L.orc_decentralize.argtypes = [vp]
    L.orc_update_global_points.argtypes = [vp]
    L.orc_get_world_points.argtypes = [vp, vp]
    L.orc_get_world_normals.argtypes = [vp, vp]
    L.orc_get_dense_tforms.argtypes = [vp, vp, vp, vp]
    L.orc_build_sets.argtypes = [vp, P(OrcSettings)]
    L.orc_sets_counts.argtypes = [vp, P(i64), P(i64), P(i64), P(i64)]
    L.orc_sets_get.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.orc_lattice.argtypes = [vp, i64, f32, vp, vp, vp, P(i32), P(i64)]
    L.orc_cost.argtypes = [vp, vp, vp]
    L.orc_jacobian.argtypes = [vp, vp, vp]
    L.orc_iteration.argtypes = [vp, P(OrcSettings)]
    L.orc_last_trace.argtypes = [vp, vp, vp, vp, vp, vp, vp, P(C.c_int)]
    L.orc_last_rows.restype = i64
    L.orc_last_rows.argtypes = [vp]
    L.orc_optimize.argtypes = [vp, P(OrcSettings), P(C.c_int), P(C.c_int)]
    L.orc_time_cost_evals.restype = f64
    L.orc_time_cost_evals.argtypes = [vp, i32]
    L.orc_time_iteration.restype = f64
    L.orc_time_iteration.argtypes = [vp, P(OrcSettings), P(C.c_int)]
    L.orc_exp_so3.argtypes = [vp, vp]
    L.orc_log_so3.argtypes = [vp, vp]
    L.orc_slerp.argtypes = [vp, vp, f64, vp]
    L.orc_fh_weights.argtypes = [vp, i32, i32, vp]
    L.orc_fh_eval.restype = f64
    L.orc_fh_eval.argtypes = [vp, vp, vp, i32, f64]
    L.orc_gaussian_info.argtypes = [vp, i32, vp]
    L.orc_relative2global.argtypes = [i32, vp, vp, vp, vp]
    L.orc_global2relative.argtypes = [i32, vp, vp, vp, vp]
    L.orc_lu_inverse.argtypes = [vp, i32, vp]
    L.orc_select_static_points.restype = i64
    L.orc_select_static_points.argtypes = [vp, i64, vp, i64, vp, C.c_float, C.c_float, vp]
    L.orc_overlap.restype = C.c_float
    L.orc_overlap.argtypes = [vp, i64, vp, i64, C.c_float]
    L.orc_rand_sequence.argtypes = [C.c_uint, i64, vp]
    L.orc_grid_downsample.restype = i64
    L.orc_grid_downsample.argtypes = [vp, i64, C.c_float, C.c_uint, vp]
    L.orc_preprocess.restype = i64
    L.orc_preprocess.argtypes = [vp, i64, i32, C.c_float, C.c_float, vp, C.c_uint, vp, P(C.c_float)]
    L.orc_update_normals.argtypes = [vp, i64, vp, vp]
    _LIBS[opt] = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def window_timing(t_min, t_max, n_poses, dt_res):
    """Restatement of ContinuousTrajectory::initTraj (ContinuousTrajectory.h:301-346) for feeding the oracle."""
    from dmsa_lidar_slam_b200.synth import linspaced

    horizon = t_max - t_min + dt_res
    n_total = int(round(horizon / dt_res)) + 1
    traj_time = linspaced(n_total, 0.0, horizon)
    stamps = linspaced(n_poses, 0.0, horizon)
    param_indices = np.round(stamps / dt_res).astype(np.int32)
    return dict(t0=t_min, horizon=horizon, n_total=n_total, traj_time=traj_time, stamps=stamps, param_indices=param_indices)


def tform_ids(stamps, t0, traj_time):
    """registerPcBuffer (ContinuousTrajectory.h:251-254): lower_bound of stamp - t0, clamped."""
    idx = np.searchsorted(traj_time, np.asarray(stamps, dtype=np.float64) - t0, side="left")
    return np.minimum(idx, len(traj_time) - 1).astype(np.int32)


class OracleModel:
    """Thin OO wrapper over the oracle's C API."""

    def __init__(self, handle, kind, opt="O2"):
        self.L = lib(opt)
        self.h = handle
        self.kind = kind
        self._keep = []

    @classmethod
    def from_window(cls, win, opt="O2"):
        L = lib(opt)
        tim = window_timing(win["t_min"], win["t_max"], win["n_poses"], win["dt_res"])
        scan = np.concatenate(win["scans"])
        xyzw = np.ascontiguousarray(np.stack([scan["x"], scan["y"], scan["z"], scan["w"]], axis=1), dtype=np.float32)
        tid = tform_ids(scan["stamp"], tim["t0"], tim["traj_time"])
        ring = np.ascontiguousarray(scan["id"], dtype=np.int32)
        st = win["static"]
        sxyzw = np.ascontiguousarray(np.stack([st["x"], st["y"], st["z"], st["w"]], axis=1), dtype=np.float32)
        sring = np.ascontiguousarray(st["id"], dtype=np.int32)
        ro = np.ascontiguousarray(win["rel_orient"].T, dtype=np.float64)  # column-major 3 x n == row-major n x 3
        rt = np.ascontiguousarray(win["rel_transl"].T, dtype=np.float64)
        h = L.orc_traj_create(win["n_poses"], _p(tim["stamps"]), tim["n_total"], _p(tim["traj_time"]), win["dt_res"], _p(xyzw), _p(tid),
                              _p(ring), len(scan), _p(sxyzw), _p(sring), len(st), float(min(win["grid_sizes"])), _p(ro), _p(rt))
        m = cls(h, "traj", opt)
        m.timing = tim
        m.N = len(scan) + len(st)
        m.n_scan = len(scan)
        m.n_poses = win["n_poses"]
        m.tid = tid
        return m

    @classmethod
    def from_submap(cls, sm, opt="O2"):
        L = lib(opt)
        pts = np.concatenate(sm["clouds"])
        xyzw = np.ascontiguousarray(np.stack([pts["x"], pts["y"], pts["z"], pts["w"]], axis=1), dtype=np.float32)
        nrm = np.ascontiguousarray(np.stack([pts["nx"], pts["ny"], pts["nz"], pts["nw"]], axis=1), dtype=np.float32)
        kf = np.concatenate([np.full(len(c), k, dtype=np.int32) for k, c in enumerate(sm["clouds"])])
        ring = np.ascontiguousarray(np.concatenate(sm["rings"]), dtype=np.int32)
        ro = np.ascontiguousarray(sm["rel_orient"].T, dtype=np.float64)
        rt = np.ascontiguousarray(sm["rel_transl"].T, dtype=np.float64)
        h = L.orc_kf_create(sm["n_keyframes"], _p(xyzw), _p(nrm), _p(kf), _p(ring), len(pts), float(min(sm["grid_sizes"])), _p(ro), _p(rt))
        m = cls(h, "kf", opt)
        m.N = len(pts)
        m.n_scan = len(pts)
        m.n_poses = sm["n_keyframes"]
        return m

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- parameters / poses
    @property
    def P(self):
        return self.L.orc_num_params(self.h)

    @property
    def E(self):
        return self.L.orc_num_extra(self.h)

    def set_mode(self, mode):
        self.L.orc_set_mode(self.h, {"faithful": 0, "f64": 1}.get(mode, mode))

    def set_threads(self, t):
        self.L.orc_set_threads(self.h, int(t))

    def get_params(self):
        p = np.zeros(self.P)
        self.L.orc_get_params(self.h, _p(p))
        return p

    def set_params(self, p):
        p = c64(p)
        self.L.orc_set_params(self.h, _p(p))

    def get_poses(self):
        n = self.n_poses
        out = [np.zeros((n, 3)) for _ in range(4)]
        self.L.orc_get_poses(self.h, *[_p(o) for o in out])
        return dict(rel_orient=out[0].T.copy(), rel_transl=out[1].T.copy(), glob_orient=out[2].T.copy(), glob_transl=out[3].T.copy())

    def centralize(self):
        self.L.orc_centralize(self.h)

    def decentralize(self):
        self.L.orc_decentralize(self.h)

    def update_global_points(self):
        self.L.orc_update_global_points(self.h)

    def world_points(self):
        w = np.zeros((self.N, 4), dtype=np.float32)
        self.L.orc_get_world_points(self.h, _p(w))
        return w

    def world_normals(self):
        w = np.zeros((self.N, 4), dtype=np.float32)
        self.L.orc_get_world_normals(self.h, _p(w))
        return w

    def dense_tforms(self):
        nt = self.timing["n_total"]
        M = np.zeros((nt, 12), dtype=np.float32)
        O = np.zeros((nt, 3))
        T = np.zeros((nt, 3))
        self.L.orc_get_dense_tforms(self.h, _p(M), _p(O), _p(T))
        return M, O, T

    # ---- sets
    def build_sets(self, st):
        return self.L.orc_build_sets(self.h, C.byref(st))

    def sets(self):
        G, M, raw, mm = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self.L.orc_sets_counts(self.h, C.byref(G), C.byref(M), C.byref(raw), C.byref(mm))
        G, M = G.value, M.value
        offs = np.zeros(G + 1, dtype=np.int64)
        members = np.zeros(M, dtype=np.int32)
        info = np.zeros((G, 9), dtype=np.float32)
        w = np.zeros(G, dtype=np.float32)
        level = np.zeros(G, dtype=np.int32)
        key = np.zeros((G, 3), dtype=np.int32)
        sub = np.zeros(G, dtype=np.int32)
        self.L.orc_sets_get(self.h, _p(offs), _p(members), _p(info), _p(w), _p(level), _p(key), _p(sub))
        return dict(G=G, M=M, raw=raw.value, lattice_mismatch=mm.value, offs=offs, members=members, info=info, w=w, level=level, key=key, sub=sub)

    # ---- cost / jacobian / iteration
    def n_rows(self):
        G = C.c_int64()
        d = C.c_int64()
        self.L.orc_sets_counts(self.h, C.byref(G), C.byref(d), C.byref(d), C.byref(d))
        return G.value + self.E

    def cost(self, p=None):
        e = np.zeros(self.n_rows())
        pp = c64(p) if p is not None else None
        self.L.orc_cost(self.h, _p(pp), _p(e))
        return e

    def jacobian(self):
        R = self.n_rows()
        e0 = np.zeros(R)
        J = np.zeros((self.P, R))  # column-major R x P
        self.L.orc_jacobian(self.h, _p(e0), _p(J))
        return e0, J.T.copy()

    def iteration(self, st):
        return self.L.orc_iteration(self.h, C.byref(st))

    def last_trace(self, with_J=False):
        R = self.L.orc_last_rows(self.h)
        P = self.P
        e0 = np.zeros(R)
        J = np.zeros((P, R)) if with_J else None
        H = np.zeros((P, P))
        g = np.zeros(P)
        step = np.zeros(P)
        ls = np.zeros(9)
        bk = C.c_int()
        self.L.orc_last_trace(self.h, _p(e0), _p(J), _p(H), _p(g), _p(step), _p(ls), C.byref(bk))
        return dict(e0=e0, J=(J.T.copy() if with_J else None), H=H, g=g, step=step, ls_cost=ls, best_k=bk.value)

    def optimize(self, st):
        it, reason = C.c_int(), C.c_int()
        self.L.orc_optimize(self.h, C.byref(st), C.byref(it), C.byref(reason))
        return it.value, reason.value

    def time_cost_evals(self, n):
        return self.L.orc_time_cost_evals(self.h, int(n))

    def time_iteration(self, st):
        s = C.c_int()
        t = self.L.orc_time_iteration(self.h, C.byref(st), C.byref(s))
        return t, s.value


def select_static_points(window_xyzw, cloud_xyz_nrm, pos, max_dist_sq, radius):
    """addStaticPoints inner loop for one keyframe cloud (DmsaSlam.h:304-339): (selected uint8[n], currOverlap)."""
    L = lib()
    w = np.ascontiguousarray(window_xyzw, dtype=np.float32).reshape(-1, 4)
    c = np.ascontiguousarray(cloud_xyz_nrm, dtype=np.float32).reshape(-1, 8)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    sel = np.zeros(len(c), dtype=np.uint8)
    cnt = L.orc_select_static_points(_p(w), len(w), _p(c), len(c), _p(pos), float(max_dist_sq), float(radius), _p(sel))
    return sel, int(cnt)


def overlap(pc1_xyzw, window_xyzw, max_dist):
    """getOverlap(pc1, window, maxDist), DmsaSlam.h:377-414."""
    L = lib()
    a = np.ascontiguousarray(pc1_xyzw, dtype=np.float32).reshape(-1, 4)
    w = np.ascontiguousarray(window_xyzw, dtype=np.float32).reshape(-1, 4)
    return float(L.orc_overlap(_p(a), len(a), _p(w), len(w), float(max_dist)))


# ---- SURVEY 8(f) rank 3 (helpers.h:67-182, DmsaSlam.h:557-634) -------------------------------------------------------------
def rand_sequence(seed, n):
    out = np.zeros(int(n), dtype=np.int32)
    lib().orc_rand_sequence(int(seed) & 0xFFFFFFFF, int(n), _p(out))
    return out


def grid_downsample(xyz, grid, seed):
    """xyz: (N, 3) float32 -> picked indices in leaf order."""
    xyzw = np.ones((len(xyz), 4), dtype=np.float32)
    xyzw[:, :3] = xyz
    idx = np.zeros(len(xyz), dtype=np.int32)
    n = lib().orc_grid_downsample(_p(xyzw), len(xyz), float(grid), int(seed) & 0xFFFFFFFF, _p(idx))
    return idx[:n]


def preprocess(raw, max_num, minDistDS, min_dist, T, seed):
    """raw: PointStampId structured array; T: 4 x 4 -> (filtered records, grid size)."""
    raw = np.ascontiguousarray(raw)
    out = np.zeros(len(raw), dtype=raw.dtype)
    Tc = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(4, 4).T).ravel()  # column-major
    gs = C.c_float(0.0)
    n = lib().orc_preprocess(_p(raw), len(raw), int(max_num), float(minDistDS), float(min_dist), _p(Tc), int(seed) & 0xFFFFFFFF, _p(out), C.byref(gs))
    return out[:n], float(gs.value)


def update_normals(cloud, origin=(0.0, 0.0, 0.0)):
    """cloud: pcl::PointNormal structured array -> (cloud with normals / curvature, (n, 6) neighbour indices)."""
    c = np.ascontiguousarray(cloud).copy()
    assert c.dtype.itemsize == 48
    nn = np.zeros((len(c), 6), dtype=np.int32)
    vp_ = np.asarray(origin, dtype=np.float32)
    lib().orc_update_normals(_p(c), len(c), _p(vp_), _p(nn))
    return c, nn
