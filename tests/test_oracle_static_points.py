"""SURVEY §8(f) rank 2 - static-point selection and overlap ratio (DmsaSlam.h:264-414): the oracle's grid restatement
against a brute-force float32 replay of the reference's arithmetic (FLANN L2_Simple order, Eigen 3-vector dot order) and
against scipy's kd-tree."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import oracle_binding as ob


def l2_simple(q, pts):
    """flann::L2_Simple<float>: result = ((0 + dx*dx) + dy*dy) + dz*dz in float32."""
    d = (q[None, :3].astype(np.float32) - pts[:, :3].astype(np.float32)).astype(np.float32)
    r = (d[:, 0] * d[:, 0]).astype(np.float32)
    r = (r + (d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
    r = (r + (d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
    return r


def visible(pos, p, n):
    f = np.float32
    d = f(f(p[0] * n[0]) + f(f(p[1] * n[1]) + f(p[2] * n[2])))
    e = f(f(pos[0] * n[0]) + f(f(pos[1] * n[1]) + f(pos[2] * n[2])))
    return float(f(e - d)) >= -0.00001


def make_case(seed, n_win=4000, n_kf=1500, radius=0.3):
    rng = np.random.default_rng(seed)
    win = np.ones((n_win, 4), dtype=np.float32)
    win[:, :3] = rng.uniform(-6, 6, (n_win, 3)).astype(np.float32)
    cloud = np.zeros((n_kf, 8), dtype=np.float32)
    # a third of the keyframe points sit on top of window points (exact hits), a third at ~radius from one (boundary cases)
    src = win[rng.integers(0, n_win, n_kf), :3]
    dirs = rng.normal(size=(n_kf, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dist = np.where(rng.random(n_kf) < 0.33, 0.0, np.where(rng.random(n_kf) < 0.5, radius * (1 + 1e-6 * rng.normal(size=n_kf)), rng.uniform(0, 3 * radius, n_kf)))
    cloud[:, :3] = (src + dirs * dist[:, None]).astype(np.float32)
    cloud[:, 3] = 1.0
    nrm = rng.normal(size=(n_kf, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    cloud[:, 4:7] = nrm.astype(np.float32)
    pos = rng.uniform(-6, 6, 3).astype(np.float32)
    return win, cloud, pos


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_select_static_points_matches_bruteforce_float32(seed):
    radius = np.float32(0.3)
    max_sq = np.float32(float(radius) ** 2)  # std::pow(1.0f * minGridSize, 2) -> double, stored in a float (DmsaSlam.h:293)
    win, cloud, pos = make_case(seed)
    sel, cnt = ob.select_static_points(win, cloud, pos, max_sq, radius)
    ref = np.zeros(len(cloud), dtype=np.uint8)
    for j in range(len(cloud)):
        near = bool((l2_simple(cloud[j], win) <= max_sq).any())
        ref[j] = 1 if (near and visible(pos, cloud[j, :3], cloud[j, 4:7])) else 0
    assert np.array_equal(sel, ref)
    assert cnt == int(ref.sum()) and 0 < cnt < len(cloud)
    # kd-tree in double: same decisions away from the float rounding band around the radius
    dd, _ = cKDTree(win[:, :3].astype(np.float64)).query(cloud[:, :3].astype(np.float64))
    clear = np.abs(dd - float(radius)) > 1e-4
    vis = np.array([visible(pos, cloud[j, :3], cloud[j, 4:7]) for j in range(len(cloud))])
    assert np.array_equal(sel[clear] == 1, (dd[clear] <= float(radius)) & vis[clear])


def test_overlap_ratio_matches_bruteforce_and_edge_cases():
    radius = np.float32(0.3)
    win, cloud, _ = make_case(7, n_win=3000, n_kf=900)
    active = np.ones((len(cloud), 4), dtype=np.float32)
    active[:, :3] = cloud[:, :3]
    got = ob.overlap(active, win, radius)
    sq = np.float32(radius * radius)
    n = sum(bool((l2_simple(win[j], active) <= sq).any()) for j in range(len(win)))
    assert got == float(np.float32(n) / np.float32(len(win))) and 0.0 < got < 1.0
    assert ob.overlap(active[:0], win, radius) == 0.0 and ob.overlap(active, win[:0], radius) == 0.0  # DmsaSlam.h:380-381
    assert ob.overlap(win, win, radius) == 1.0
    far = win.copy()
    far[:, 0] += 100.0
    assert ob.overlap(far, win, radius) == 0.0
    bad = win.copy()
    bad[::7, 1] = np.nan  # non-finite reference points are never a neighbour
    assert 0.0 < ob.overlap(bad, win, radius) <= 1.0
    sel, cnt = ob.select_static_points(win[:0], cloud, np.zeros(3, np.float32), sq, radius)
    assert cnt == 0 and not sel.any()
