"""SURVEY 8(f) rank 3 on the GPU, through the C-ABI, against the oracle: randomGridDownsampling (helpers.h:67-182),
preProcess (DmsaSlam.h:570-634), updateNormals (DmsaSlam.h:557-568)."""
import numpy as np
import pytest

import oracle_binding as ob
from dmsa_lidar_slam_b200 import ContinuousTrajectory, PreProcessor, PreprocessConfig, synth
from dmsa_lidar_slam_b200.synth import POINT_NORMAL

pytestmark = pytest.mark.gpu


def _scan(n, seed=3, sensor="cfg1"):
    win = synth.make_sliding_window(n_scans=1, sensor=sensor, n_static=0, n_poses=4, seed=seed)
    return win["scans"][0][:n]


def _xyz(c):
    return np.stack([c["x"], c["y"], c["z"]], 1)


def test_random_grid_downsampling_picks_the_oracles_points():
    pre = PreProcessor()
    sc = _scan(20000)
    for grid, seed in ((0.4, 1), (0.3, 1700000000), (0.15, 2**31 + 7)):
        filt, idx = pre.randomGridDownsampling(sc, grid, seed)
        want = ob.grid_downsample(_xyz(sc), grid, seed)
        assert np.array_equal(idx, want) and len(idx) > 1000
        assert np.array_equal(filt["stamp"], sc["stamp"][want])
    # pcl::PointNormal records (48-byte stride), non-finite points (skipped by the octree), a full OS1-64 scan
    big = _scan(65536, seed=8, sensor="os1-64")
    pn = np.zeros(len(big), dtype=POINT_NORMAL)
    pn["x"], pn["y"], pn["z"], pn["w"] = big["x"], big["y"], big["z"], 1.0
    pn["x"][17] = np.nan
    pn["z"][4000] = np.inf
    filt, idx = pre.randomGridDownsampling(pn, 0.2, 99)
    want = ob.grid_downsample(_xyz(pn), 0.2, 99)
    assert np.array_equal(idx, want) and 17 not in idx and 4000 not in idx
    # degenerate inputs
    assert len(pre.randomGridDownsampling(sc[:0], 0.3, 1)[1]) == 0
    one = pre.randomGridDownsampling(sc[:1], 0.3, 1)[1]
    assert np.array_equal(one, [0])


def test_downsampling_of_the_staged_window_cloud():
    """addNewKeyframeToMap draws the keyframe cloud from trajIn.globalPoints (DmsaSlam.h:506)."""
    import ctypes as C

    win = synth.make_config("cfg1")
    traj = ContinuousTrajectory.from_window(win)
    traj.updateGlobalPoints()
    world = traj.globalPoints()
    idx = np.zeros(len(world), dtype=np.int32)
    n_out = C.c_int64(0)
    traj.ctx._ck(traj.L.dmsa_b200_downsample_global_points(traj.h, C.c_float(0.3), 5, idx.ctypes.data_as(C.c_void_p), C.byref(n_out)))
    want = ob.grid_downsample(world[:, :3], 0.3, 5)
    assert np.array_equal(idx[: n_out.value], want)


def test_preprocess_equals_the_oracle_record_for_record():
    pre = PreProcessor()
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = synth.Rot.from_rotvec([0.02, -0.01, 0.3]).as_matrix().astype(np.float32)
    T[:3, 3] = [0.1, -0.2, 0.05]
    sc = _scan(65536, seed=4, sensor="os1-64")
    used = set()
    for max_num, dds, dmin in ((3000, 30.0, 0.0), (1000, 10.0, 1.0), (6000, 8.0, 0.5), (10**6, 5.0, 0.0)):
        cfg = PreprocessConfig(max_num, dds, dmin, T)
        out, gs = pre.preProcess(sc, cfg, seed=21)
        want, gs_o = ob.preprocess(sc, max_num, dds, dmin, T, seed=21)
        assert gs == pytest.approx(gs_o)
        used.add(round(gs, 2))
        assert len(out) == len(want) and len(out) > 100
        assert out.tobytes() == want.tobytes()  # transformed xyz bit for bit, w = 1, stamp / id / isStatic copied
    assert len(used) >= 2, "the cases must exercise the adaptive grid"
    # empty scan
    out, gs = pre.preProcess(sc[:0], PreprocessConfig(), seed=1)
    assert len(out) == 0


def _keyframe_cloud(n, seed, grid=0.3, sensor="os1-64", raw=65536):
    sc = _scan(raw, seed, sensor)
    idx = ob.grid_downsample(_xyz(sc), grid, seed)[:n]
    c = np.zeros(len(idx), dtype=POINT_NORMAL)
    c["x"], c["y"], c["z"], c["w"] = sc["x"][idx], sc["y"][idx], sc["z"][idx], 1.0
    return c


@pytest.mark.parametrize("n,cell", [(3000, 0.3), (20000, 0.3), (20000, 1.0), (20000, 0.05)])
def test_update_normals_neighbour_sets_bit_exact_and_normals_within_float_noise(n, cell):
    """k = 6 neighbour index lists identical to the exhaustive search (ascending (distance, index): FLANN's result order
    for distinct distances), whatever the cell size of the search grid; normals / curvature equal to the oracle's PCL
    restatement up to the last-bit differences of atan2f / cosf / sinf between CUDA and glibc."""
    pre = PreProcessor()
    c = _keyframe_cloud(n, seed=6)
    got, nn = pre.updateNormals(c, origin=(0.5, -1.0, 0.2), cell_size=cell, with_neighbours=True)
    want, nn_o = ob.update_normals(c, origin=(0.5, -1.0, 0.2))
    assert np.array_equal(nn, nn_o)
    a = np.stack([got["nx"], got["ny"], got["nz"]], 1)
    b = np.stack([want["nx"], want["ny"], want["nz"]], 1)
    assert np.isfinite(a).all()
    err = np.abs(a - b).max(1)
    # tolerance: 1e-4 on the unit normals for >= 99.9 % of the points (a 1-ulp change of theta moves an ill-conditioned
    # eigenvector visibly; sign flips are included in the count)
    assert (err < 1e-4).mean() > 0.999, (err < 1e-4).mean()
    assert np.median(err) < 1e-6
    ok = err < 1e-4
    assert np.allclose(got["curvature"][ok], want["curvature"][ok], rtol=2e-3, atol=1e-6)
    assert np.array_equal(_xyz(got), _xyz(c)) and (got["nw"] == 0).all()


def test_update_normals_edge_cases():
    pre = PreProcessor()
    # fewer than 3 points: PCL writes NaN normals; fewer than 6: all of them are the neighbourhood
    c = _keyframe_cloud(5, seed=2)
    got, nn = pre.updateNormals(c[:2], with_neighbours=True)
    assert np.isnan(got["nx"]).all() and (nn[:, 2:] == -1).all()
    got, nn = pre.updateNormals(c, with_neighbours=True)
    want, nn_o = ob.update_normals(c)
    assert np.array_equal(nn, nn_o) and (nn[:, 5] == -1).all()
    assert np.allclose(np.stack([got["nx"], got["ny"], got["nz"]], 1), np.stack([want["nx"], want["ny"], want["nz"]], 1), atol=1e-4, equal_nan=True)
    # an isolated far point (exhaustive fall-back of the shell search) and a non-finite point
    c = _keyframe_cloud(4000, seed=9)
    c["x"][100], c["y"][100], c["z"][100] = 900.0, -700.0, 300.0
    c["y"][200] = np.nan
    got, nn = pre.updateNormals(c, cell_size=0.3, with_neighbours=True)
    want, nn_o = ob.update_normals(c)
    assert np.array_equal(nn, nn_o)
    assert np.isnan(got["nx"][200]) and (nn[200] == -1).all() and 200 not in nn[np.arange(len(c)) != 200]
