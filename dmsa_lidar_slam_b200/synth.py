"""Deterministic synthetic LiDAR windows / keyframe submaps of the BASELINE.json shapes (`synth_v1`).

Not part of the reference: the reference is validated on recorded datasets only
(README.md:93-95) and its pre-processing is non-deterministic (helpers.h:85), so the
hot path is exercised at the `optimizeSet` boundary on synthetic inputs.

Scene: closed box room 40 x 30 x 6 m plus 8 axis-aligned 1 x 1 x 6 m pillars; every ray hits
a surface.  Sensor: H x W rays, 10 Hz, per-point stamps, true motion distortion, 1 cm range
noise.  The layouts produced are exactly the reference's input layouts: `PointStampId`
(PointStampId.h:33-45, 32 B) for the sliding window, `pcl::PointNormal` (48 B) for keyframes.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as Rot

# PointStampId.h:33-45  (PCL_ADD_POINT4D; double stamp; int id; int isStatic) EIGEN_ALIGN16
POINT_STAMP_ID = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("stamp", "<f8"), ("id", "<i4"), ("isStatic", "<i4")], align=True
)
assert POINT_STAMP_ID.itemsize == 32
# pcl::PointNormal: data[4], data_n[4], curvature + 3 pad floats = 48 B
POINT_NORMAL = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("nw", "<f4"),
     ("curvature", "<f4"), ("pad", "<f4", (3,))], align=True
)
assert POINT_NORMAL.itemsize == 48

ROOM = np.array([40.0, 30.0, 6.0])
PILLARS_XY = np.array([[8, 6], [8, 22], [16, 10], [16, 20], [24, 6], [24, 22], [32, 10], [32, 20]], dtype=np.float64)
SCAN_PERIOD = 0.1

# named sensor shapes (BASELINE.json configs)
SENSORS = {
    "cfg1": dict(H=20, W=1000, fov_deg=15.0),
    "os1-64": dict(H=64, W=1024, fov_deg=16.6),
    "os0-128": dict(H=128, W=1024, fov_deg=45.0),
    "stress": dict(H=128, W=2048, fov_deg=45.0),
    "tiny": dict(H=8, W=256, fov_deg=15.0),
}


def truth_pose(t):
    """Ground-truth sensor motion p(t), R(t)."""
    t = np.asarray(t, dtype=np.float64)
    p = np.stack([2.0 + 1.0 * t, 15.0 + 0.5 * np.sin(0.8 * t), 1.5 + 0.05 * np.sin(2.0 * t)], axis=-1)
    yaw, roll, pitch = 0.3 * t, 0.05 * np.sin(1.3 * t), 0.03 * np.cos(0.9 * t)
    R = Rot.from_euler("ZYX", np.stack([yaw, pitch, roll], axis=-1))
    return p, R


def _raycast(o, d):
    """Distance along unit rays (o: n x 3 origins inside the room, d: n x 3) to the first surface; also returns the normal."""
    n = o.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # room (from inside): exit distance per axis
        t_ax = np.where(d > 0, (ROOM - o) * inv, np.where(d < 0, (0.0 - o) * inv, np.inf))
    axis = np.argmin(t_ax, axis=1)
    t_hit = t_ax[np.arange(n), axis]
    normal = np.zeros((n, 3))
    normal[np.arange(n), axis] = -np.sign(d[np.arange(n), axis])
    for px, py in PILLARS_XY:
        lo = np.array([px - 0.5, py - 0.5, 0.0])
        hi = np.array([px + 0.5, py + 0.5, ROOM[2]])
        with np.errstate(invalid="ignore"):
            t1 = (lo - o) * inv
            t2 = (hi - o) * inv
        tn = np.minimum(t1, t2)
        tf = np.maximum(t1, t2)
        tn = np.where(np.isnan(tn), -np.inf, tn)
        tf = np.where(np.isnan(tf), np.inf, tf)
        ax_n = np.argmax(tn, axis=1)
        t_near = tn[np.arange(n), ax_n]
        t_far = tf.min(axis=1)
        hit = (t_near <= t_far) & (t_near > 1e-9) & (t_near < t_hit)
        t_hit = np.where(hit, t_near, t_hit)
        nn = np.zeros((n, 3))
        nn[np.arange(n), ax_n] = -np.sign(d[np.arange(n), ax_n])
        normal = np.where(hit[:, None], nn, normal)
    return t_hit, normal


def _ray_dirs(H, W, fov_deg):
    el = np.deg2rad(np.linspace(-fov_deg, fov_deg, H))
    az = 2.0 * np.pi * np.arange(W) / W
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (H, W))], axis=-1)
    return d.reshape(-1, 3)  # ring-major: index = row*W + col


def _sample_static(n_static, center, rng, H):
    """Exactly n_static samples on scene surfaces within 30 m of `center`, world frame, 1 cm normal noise."""
    faces = []  # (origin, u, v, normal)
    X, Y, Z = ROOM
    faces += [
        (np.array([0, 0, 0.0]), np.array([X, 0, 0.0]), np.array([0, Y, 0.0]), np.array([0, 0, 1.0])),
        (np.array([0, 0, Z]), np.array([X, 0, 0.0]), np.array([0, Y, 0.0]), np.array([0, 0, -1.0])),
        (np.array([0, 0, 0.0]), np.array([X, 0, 0.0]), np.array([0, 0, Z]), np.array([0, 1.0, 0])),
        (np.array([0, Y, 0.0]), np.array([X, 0, 0.0]), np.array([0, 0, Z]), np.array([0, -1.0, 0])),
        (np.array([0, 0, 0.0]), np.array([0, Y, 0.0]), np.array([0, 0, Z]), np.array([1.0, 0, 0])),
        (np.array([X, 0, 0.0]), np.array([0, Y, 0.0]), np.array([0, 0, Z]), np.array([-1.0, 0, 0])),
    ]
    for px, py in PILLARS_XY:
        x0, x1, y0, y1 = px - 0.5, px + 0.5, py - 0.5, py + 0.5
        faces += [
            (np.array([x0, y0, 0.0]), np.array([1.0, 0, 0]), np.array([0, 0, Z]), np.array([0, -1.0, 0])),
            (np.array([x0, y1, 0.0]), np.array([1.0, 0, 0]), np.array([0, 0, Z]), np.array([0, 1.0, 0])),
            (np.array([x0, y0, 0.0]), np.array([0, 1.0, 0]), np.array([0, 0, Z]), np.array([-1.0, 0, 0])),
            (np.array([x1, y0, 0.0]), np.array([0, 1.0, 0]), np.array([0, 0, Z]), np.array([1.0, 0, 0])),
        ]
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v, _ in faces])
    prob = areas / areas.sum()
    out = np.zeros((0, 3))
    while out.shape[0] < n_static:
        m = max(1024, int(1.6 * (n_static - out.shape[0])))
        f = rng.choice(len(faces), size=m, p=prob)
        a, b = rng.random(m), rng.random(m)
        O = np.stack([faces[i][0] for i in f])
        U = np.stack([faces[i][1] for i in f])
        V = np.stack([faces[i][2] for i in f])
        Nn = np.stack([faces[i][3] for i in f])
        P = O + a[:, None] * U + b[:, None] * V + Nn * rng.normal(0.0, 0.01, m)[:, None]
        keep = np.linalg.norm(P - center, axis=1) <= 30.0
        out = np.concatenate([out, P[keep]], axis=0)
    out = out[:n_static]
    pts = np.zeros(n_static, dtype=POINT_STAMP_ID)
    pts["x"], pts["y"], pts["z"] = out[:, 0].astype(np.float32), out[:, 1].astype(np.float32), out[:, 2].astype(np.float32)
    pts["w"] = 1.0
    pts["stamp"] = -1000.0  # ContinuousTrajectory.h:168
    pts["id"] = rng.integers(0, H, n_static, dtype=np.int32)
    pts["isStatic"] = 1
    return pts


def linspaced(n, lo, hi):
    """Eigen VectorXd::LinSpaced(n, lo, hi): lo + i*step, last element == hi."""
    if n == 1:
        return np.array([hi], dtype=np.float64)
    step = (hi - lo) / (n - 1)
    v = lo + np.arange(n, dtype=np.float64) * step
    v[-1] = hi
    return v


def make_sliding_window(n_scans=1, sensor="cfg1", n_static=5000, n_poses=4, seed=0, dt_res=1e-3, grid_size=0.3,
                        sigma_t=0.02, sigma_r=0.005):
    """One sliding window: `n_scans` motion-distorted scans + static map points + perturbed control poses.

    Returns a dict with the reference-shaped inputs:
      scans      list of POINT_STAMP_ID arrays (IMU frame, chronological)
      static     POINT_STAMP_ID array (world frame)
      rel_orient / rel_transl  3 x n_poses float64, column per pose (Poses.h:19-20), relative chain
      t_min, t_max, dt_res, n_poses, grid_sizes (per scan)
    """
    sp = SENSORS[sensor]
    H, W, fov = sp["H"], sp["W"], sp["fov_deg"]
    rng_noise = np.random.default_rng([0xD35A, seed, 1])
    rng_pert = np.random.default_rng([0xD35A, seed, 2])
    rng_static = np.random.default_rng([0xD35A, seed, 3])
    dirs = _ray_dirs(H, W, fov)
    col = np.tile(np.arange(W), H)
    row = np.repeat(np.arange(H), W)
    scans = []
    for s in range(n_scans):
        stamps = s * SCAN_PERIOD + col * (SCAN_PERIOD / W)
        p, R = truth_pose(stamps)
        dw = R.apply(dirs)
        rng_, _ = _raycast(p, dw)
        rng_ = rng_ + rng_noise.normal(0.0, 0.01, rng_.shape)
        loc = (dirs * rng_[:, None]).astype(np.float32)
        pc = np.zeros(H * W, dtype=POINT_STAMP_ID)
        pc["x"], pc["y"], pc["z"], pc["w"] = loc[:, 0], loc[:, 1], loc[:, 2], 1.0
        pc["stamp"] = stamps
        pc["id"] = row.astype(np.int32)
        pc["isStatic"] = 0
        scans.append(pc)
    t_min = 0.0
    t_max = float(max(sc["stamp"].max() for sc in scans))
    horizon = t_max - t_min + dt_res  # ContinuousTrajectory.h:309
    ctrl = linspaced(n_poses, 0.0, horizon)  # :332
    p, R = truth_pose(t_min + ctrl)
    Rm = R.as_matrix()
    rel_o = np.zeros((3, n_poses))
    rel_t = np.zeros((3, n_poses))
    rel_o[:, 0] = R[0].as_rotvec()
    rel_t[:, 0] = p[0]
    for k in range(1, n_poses):
        rel_o[:, k] = Rot.from_matrix(Rm[k - 1].T @ Rm[k]).as_rotvec()
        rel_t[:, k] = Rm[k - 1].T @ (p[k] - p[k - 1])
    rel_o[:, 1:] += rng_pert.normal(0.0, sigma_r, (3, n_poses - 1))
    rel_t[:, 1:] += rng_pert.normal(0.0, sigma_t, (3, n_poses - 1))
    p0, _ = truth_pose(0.0)
    static = _sample_static(n_static, p0, rng_static, H) if n_static > 0 else np.zeros(0, dtype=POINT_STAMP_ID)
    return dict(scans=scans, static=static, rel_orient=rel_o, rel_transl=rel_t, t_min=t_min, t_max=t_max, dt_res=dt_res,
                n_poses=n_poses, grid_sizes=[grid_size] * n_scans, H=H, W=W)


def make_keyframe_submap(n_keyframes=4, n_points=2000, seed=0, spacing=2.0, grid_size=0.3, sigma_t=0.02, sigma_r=0.005):
    """A keyframe submap (MapManagement/KeyframeData.h): per keyframe a local PointNormal cloud + ring ids.

    Keyframes sit every `spacing` metres along the room's long axis; points are ray-cast from a static
    sensor with exactly `n_points` rays; normals are the true surface normals (they already face the sensor).
    """
    rng = np.random.default_rng([0xD35A, seed, 4])
    H = 32
    Wc = int(np.ceil(n_points / H))
    dirs_all = _ray_dirs(H, Wc, 30.0)[:n_points]
    ring_all = np.repeat(np.arange(H), Wc)[:n_points].astype(np.int32)
    clouds, rings = [], []
    P = np.zeros((n_keyframes, 3))
    Rm = np.zeros((n_keyframes, 3, 3))
    for k in range(n_keyframes):
        s = k * spacing
        pos = np.array([3.0 + (s % 34.0), 15.0 + 2.0 * np.sin(0.3 * s), 1.5])
        R = Rot.from_euler("ZYX", [0.15 * s, 0.02 * np.sin(s), 0.03 * np.cos(s)])
        P[k], Rm[k] = pos, R.as_matrix()
        dw = R.apply(dirs_all)
        r, nrm = _raycast(np.broadcast_to(pos, dw.shape).copy(), dw)
        r = r + rng.normal(0.0, 0.01, r.shape)
        loc = (dirs_all * r[:, None]).astype(np.float32)
        nloc = (R.inv().apply(nrm)).astype(np.float32)
        pc = np.zeros(n_points, dtype=POINT_NORMAL)
        pc["x"], pc["y"], pc["z"], pc["w"] = loc[:, 0], loc[:, 1], loc[:, 2], 1.0
        pc["nx"], pc["ny"], pc["nz"], pc["nw"] = nloc[:, 0], nloc[:, 1], nloc[:, 2], 0.0
        clouds.append(pc)
        rings.append(ring_all.copy())
    rel_o = np.zeros((3, n_keyframes))
    rel_t = np.zeros((3, n_keyframes))
    rel_o[:, 0] = Rot.from_matrix(Rm[0]).as_rotvec()
    rel_t[:, 0] = P[0]
    for k in range(1, n_keyframes):
        rel_o[:, k] = Rot.from_matrix(Rm[k - 1].T @ Rm[k]).as_rotvec()
        rel_t[:, k] = Rm[k - 1].T @ (P[k] - P[k - 1])
    rel_o[:, 1:] += rng.normal(0.0, sigma_r, (3, n_keyframes - 1))
    rel_t[:, 1:] += rng.normal(0.0, sigma_t, (3, n_keyframes - 1))
    return dict(clouds=clouds, rings=rings, rel_orient=rel_o, rel_transl=rel_t, grid_sizes=[grid_size] * n_keyframes,
                n_keyframes=n_keyframes)


# BASELINE.json configs -> generator arguments
CONFIGS = {
    "cfg1": dict(n_scans=1, sensor="cfg1", n_static=5000, n_poses=4),
    "cfg2": dict(n_scans=10, sensor="os1-64", n_static=50000, n_poses=20),
    "cfg3": dict(n_scans=10, sensor="os0-128", n_static=200000, n_poses=20),
    "cfg5": dict(n_scans=20, sensor="stress", n_static=1000000, n_poses=40),
    "tiny": dict(n_scans=2, sensor="tiny", n_static=500, n_poses=3),
}


def make_config(name, seed=None):
    kw = dict(CONFIGS[name])
    if seed is None:
        seed = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg5": 5, "tiny": 0}[name]
    return make_sliding_window(seed=seed, **kw)
