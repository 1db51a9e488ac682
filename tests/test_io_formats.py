"""the data formats either side of the path (src/dmsa_slam_ros.cpp:286-291, 372-486; OutputManagement.h:80-96)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

from dmsa_lidar_slam_b200 import format_tum_pose, pc2_layout_for_sensor, save_pcd_ascii
from dmsa_lidar_slam_b200.synth import POINT_NORMAL

# (numpy record layout of one PointCloud2 point, the reference's field index of every member) per sensor type
SENSORS = {
    "hesai": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("timestamp", "<f8"), ("ring", "<u2")],
    "ouster": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("t", "<u4"), ("reflectivity", "<u2"), ("ring", "u1"), ("ambient", "<u2"), ("range", "<u4")],
    "robosense": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("ring", "<u2"), ("timestamp", "<f8")],
    "velodyne": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("ring", "<u2"), ("time", "<f4")],
    "livoxXYZRTLT_s": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("tag", "u1"), ("line", "u1"), ("timestamp", "<f8")],
    "livoxXYZRTLT_ns": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("tag", "u1"), ("line", "u1"), ("timestamp", "<f8")],
    "sick": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4"), ("range", "<f4"), ("azimuth", "<f4"), ("elevation", "<f4"), ("refl", "<f4"), ("t", "<f4"),
             ("a", "<u2"), ("b", "u1"), ("layer", "i1")],
    "unknown": [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4")],
}


def make_message(sensor, n, seed=0):
    dt = np.dtype(SENSORS[sensor])  # packed: fields at odd offsets (unaligned 8-byte stamps included)
    rng = np.random.default_rng(seed)
    m = np.zeros(n, dtype=dt)
    for name in dt.names:
        kind = dt[name].kind
        if kind == "f":
            m[name] = rng.uniform(0, 100, n) if name not in ("timestamp",) else rng.uniform(1.7e9, 1.7e9 + 0.1, n) * (1e9 if sensor.endswith("_ns") else 1.0)
        elif kind == "u":
            m[name] = rng.integers(0, min(np.iinfo(dt[name]).max, 10**8), n)
        else:
            m[name] = rng.integers(-100, 100, n)
    offsets = [dt.fields[name][1] for name in dt.names]
    return m, offsets, dt.itemsize


def expected(sensor, m, stamp_msg, delta_t):
    n = len(m)
    k = np.arange(n)
    if sensor in ("hesai", "robosense", "livoxXYZRTLT_s"):
        stamp = m["timestamp"].astype(np.float64)
    elif sensor == "ouster":
        stamp = stamp_msg + 1e-9 * m["t"].astype(np.float64)
    elif sensor == "velodyne":
        stamp = stamp_msg + m["time"].astype(np.float64)
    elif sensor == "sick":
        stamp = stamp_msg + m["t"].astype(np.float64)
    elif sensor == "livoxXYZRTLT_ns":
        stamp = 1e-9 * m["timestamp"].astype(np.float64)
    else:
        stamp = stamp_msg + delta_t * k.astype(np.float64) / float(n)
    ring = {"hesai": "ring", "ouster": "ring", "robosense": "ring", "velodyne": "ring", "sick": "layer"}.get(sensor)
    ids = m[ring].astype(np.int32) if ring else (k % 1000).astype(np.int32)
    return stamp, ids


def test_sensor_layouts_follow_the_reference_field_indices():
    m, off, step = make_message("velodyne", 4)
    L = pc2_layout_for_sensor("velodyne", off, step)
    assert (L.x_offset, L.y_offset, L.z_offset) == (0, 4, 8) and L.ring_offset == off[4] and L.stamp_offset == off[5] and L.point_step == step
    with pytest.raises(Exception):
        pc2_layout_for_sensor("no such sensor", off, step)
    with pytest.raises(Exception):
        pc2_layout_for_sensor("sick", off, step)  # needs fields[8] and fields[11]


@pytest.mark.gpu
@pytest.mark.parametrize("sensor", sorted(SENSORS))
def test_pointcloud2_decode_matches_the_reference_loop(sensor):
    from dmsa_lidar_slam_b200 import decode_pointcloud2
    from dmsa_lidar_slam_b200.api import _Context

    ctx = _Context()
    m, off, step = make_message(sensor, 70001, seed=3)
    L = pc2_layout_for_sensor(sensor, off, step)
    out = decode_pointcloud2(ctx, m.tobytes(), len(m), L, stamp_msg=1700000000.25, delta_t=0.1)
    stamp, ids = expected(sensor, m, 1700000000.25, 0.1)
    assert np.array_equal(out["x"], m["x"]) and np.array_equal(out["y"], m["y"]) and np.array_equal(out["z"], m["z"])
    assert (out["w"] == 1.0).all() and (out["isStatic"] == 0).all()
    assert np.array_equal(out["stamp"], stamp) and np.array_equal(out["id"], ids)
    assert len(decode_pointcloud2(ctx, b"", 0, L, 0.0)) == 0


def test_tum_pose_line():
    rng = np.random.default_rng(1)
    for _ in range(200):
        aa = rng.normal(0, 1.5, 3)
        if rng.random() < 0.2:  # rotations near pi: the branch of Eigen's matrix -> quaternion conversion without a positive trace
            aa = aa / np.linalg.norm(aa) * (np.pi - 10.0 ** rng.uniform(-6, -1))
        pos = rng.normal(0, 100, 3)
        stamp = 1.7e9 + rng.uniform(0, 1000)
        f = [float(v) for v in format_tum_pose(stamp, pos, aa).split()]
        assert len(f) == 8 and abs(f[0] - stamp) <= 5e-7 and np.allclose(f[1:4], pos, atol=5.1e-6)
        q = Rot.from_rotvec(aa).as_quat()  # x y z w
        got = np.array(f[4:8])
        assert min(np.abs(got - q).max(), np.abs(got + q).max()) <= 1.1e-6
    assert format_tum_pose(12.5, [1, 2, 3], [0, 0, 0]) == "12.500000 1.00000 2.00000 3.00000 0.000000 0.000000 0.000000 1.000000\n"


def test_pcd_ascii_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    c = np.zeros(257, dtype=POINT_NORMAL)
    for k in ("x", "y", "z", "nx", "ny", "nz", "curvature"):
        c[k] = rng.normal(0, 30, len(c)).astype(np.float32)
    c["nx"][5] = np.nan
    path = tmp_path / "PointCloud.pcd"
    save_pcd_ascii(path, c)
    lines = open(path).read().splitlines()
    assert lines[0].startswith("# .PCD v0.7") and "FIELDS x y z normal_x normal_y normal_z curvature" in lines and f"POINTS {len(c)}" in lines
    body = lines[lines.index("DATA ascii") + 1:]
    assert len(body) == len(c)
    got = np.array([[float(v) for v in ln.split()] for ln in body], dtype=np.float32)
    want = np.stack([c[k] for k in ("x", "y", "z", "nx", "ny", "nz", "curvature")], 1)
    # PCL's default precision (8 significant digits) does not always round-trip a float: equal to 1 part in 1e7
    assert np.allclose(got, want, rtol=1e-7, atol=0, equal_nan=True) and (got == want).mean() > 0.7
