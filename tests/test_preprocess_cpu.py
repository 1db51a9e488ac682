"""SURVEY 8(f) rank 3 on the CPU: pins of the oracle's restatements (randomGridDownsampling, preProcess, updateNormals)
against independent implementations, and the product's glibc rand() restatement against this machine's libc."""
import ctypes
import ctypes.util

import numpy as np
import pytest

import oracle_binding as ob
from dmsa_lidar_slam_b200 import rand_sequence, synth
from dmsa_lidar_slam_b200.synth import POINT_NORMAL, POINT_STAMP_ID


def _scan(n=20000, seed=3):
    win = synth.make_sliding_window(n_scans=1, sensor="cfg1", n_static=0, n_poses=4, seed=seed)
    return win["scans"][0][:n]


def test_rand_sequence_is_glibc_rand():
    """helpers.h:87-94 draws with libc rand(); the library generates the same sequence itself (TYPE_3 additive feedback
    generator, stdlib/random_r.c) so that the draw does not depend on hidden process state."""
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.rand.restype = ctypes.c_int
    for seed in (0, 1, 42, 1700000000, 2**31 + 5, 2**32 - 1):
        libc.srand(ctypes.c_uint(seed))
        want = np.array([libc.rand() for _ in range(2000)], dtype=np.int32)
        assert np.array_equal(rand_sequence(seed, 2000), want), seed
        assert np.array_equal(ob.rand_sequence(seed, 2000), want), seed


def test_grid_downsampling_takes_one_member_of_every_voxel():
    sc = _scan()
    xyz = np.stack([sc["x"], sc["y"], sc["z"]], 1)
    for grid in (0.4, 0.15):
        idx = ob.grid_downsample(xyz, grid, seed=7)
        # independent voxelisation on PCL's lattice: anchor = first point - grid (adoptBoundingBoxToPoint + getKeyBitSize)
        mn0 = xyz[0].astype(np.float64) - np.float64(np.float32(grid))
        key = np.floor((xyz.astype(np.float64) - mn0) / np.float64(np.float32(grid))).astype(np.int64)
        uniq = np.unique(key, axis=0)
        assert len(idx) == len(uniq) and len(np.unique(idx)) == len(idx)
        assert len(np.unique(key[idx], axis=0)) == len(uniq)  # one pick per voxel
        # the draw: member int(r * (n - 1)) of the leaf, members in ascending index
        r = ob.rand_sequence(7, len(idx)).astype(np.float64) / 2147483647.0
        inv = {tuple(k): np.flatnonzero((key == k).all(1)) for k in key[idx[:50]]}
        for c in range(50):
            mem = inv[tuple(key[idx[c]])]
            assert idx[c] == mem[int(r[c] * (len(mem) - 1))]
    # a different seed changes the draw, not the leaves
    a, b = ob.grid_downsample(xyz, 0.4, 1), ob.grid_downsample(xyz, 0.4, 2)
    assert len(a) == len(b) and (a != b).any()


def test_preprocess_matches_a_numpy_restatement():
    sc = _scan(60000)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = synth.Rot.from_rotvec([0.02, -0.01, 0.3]).as_matrix().astype(np.float32)
    T[:3, 3] = [0.1, -0.2, 0.05]
    for max_num, dds, dmin in ((3000, 30.0, 0.0), (1000, 10.0, 1.0), (200000, 5.0, 0.5)):
        out, gs = ob.preprocess(sc, max_num, dds, dmin, T, seed=11)
        # numpy restatement on top of the oracle's own downsampling
        xyz = np.stack([sc["x"], sc["y"], sc["z"]], 1)
        picked, used = None, None
        for g in (0.4, 0.3, 0.2, 0.15):
            if picked is None or len(picked) < max_num:
                picked, used = ob.grid_downsample(xyz, g, 11), g
        assert gs == pytest.approx(used)
        p = xyz[picked]
        rng = np.sqrt(p[:, 0] * p[:, 0] + (p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2]))
        thr = max(np.sort(rng)[min(max_num, len(rng) - 1)], np.float32(dds))
        keep = (rng < thr) & (rng > np.float32(dmin))
        q = p[keep]
        Tc = T.T.ravel()  # column-major
        want = np.stack([q[:, 0] * Tc[r] + (q[:, 1] * Tc[4 + r] + (q[:, 2] * Tc[8 + r] + Tc[12 + r])) for r in range(3)], 1).astype(np.float32)
        assert len(out) == keep.sum() and len(out) > 100
        assert np.array_equal(np.stack([out["x"], out["y"], out["z"]], 1), want)
        assert (out["w"] == 1.0).all()
        assert np.array_equal(out["stamp"], sc["stamp"][picked][keep]) and np.array_equal(out["id"], sc["id"][picked][keep])


def _keyframe_cloud(n=6000, seed=5, grid=0.3):
    sc = _scan(40000, seed)
    xyz = np.stack([sc["x"], sc["y"], sc["z"]], 1)
    idx = ob.grid_downsample(xyz, grid, seed)[:n]
    c = np.zeros(len(idx), dtype=POINT_NORMAL)
    c["x"], c["y"], c["z"], c["w"] = xyz[idx, 0], xyz[idx, 1], xyz[idx, 2], 1.0
    return c


def test_oracle_knn_matches_scipy_kdtree_and_normals_match_an_eigen_decomposition():
    from scipy.spatial import cKDTree

    c = _keyframe_cloud()
    xyz = np.stack([c["x"], c["y"], c["z"]], 1)
    out, nn = ob.update_normals(c, origin=(0, 0, 0))
    d, i = cKDTree(xyz.astype(np.float64)).query(xyz.astype(np.float64), k=6)
    # same neighbour SETS wherever the 6th and 7th distances are not within float rounding of each other
    d7 = cKDTree(xyz.astype(np.float64)).query(xyz.astype(np.float64), k=7)[0][:, 6]
    clear = (d7 - d[:, 5]) > 1e-5
    assert clear.mean() > 0.99
    assert all(set(nn[j]) == set(i[j]) for j in np.flatnonzero(clear))
    assert (nn[:, 0] == np.arange(len(c))).all()  # the point itself comes first (distance 0)
    # normals: the eigenvector of the smallest eigenvalue of the neighbourhood covariance, oriented towards the view point
    nrm = np.stack([out["nx"], out["ny"], out["nz"]], 1)
    good, agree = 0, 0
    for j in range(0, len(c), 7):
        P = xyz[nn[j]].astype(np.float64)
        w, v = np.linalg.eigh(np.cov(P.T, bias=True))
        if w[1] < 50 * max(w[0], 1e-9):  # no clear plane: the smallest eigenvector is ill-defined
            continue
        good += 1
        # PCL's one-pass float covariance loses ~1e-4 of x^2 ~ 1e3 m^2: compare directions loosely, orientation strictly
        if abs(np.dot(v[:, 0], nrm[j])) > 0.95:
            agree += 1
        assert np.dot(nrm[j], -xyz[j]) >= -1e-6
    assert good > 200 and agree / good > 0.97
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
