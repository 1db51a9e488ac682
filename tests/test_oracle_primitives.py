"""Pins the oracle's restatements of third-party arithmetic (Eigen / Boost / PCL — absent from this image) against
independent scipy / numpy / brute-force restatements.  CPU only."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial.transform import Rotation as Rot
from scipy.spatial.transform import Slerp

import oracle_binding as ob

L = ob.lib()
_p = ob._p
rng = np.random.default_rng(1234)


def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def o_exp(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    R = np.zeros(9)
    L.orc_exp_so3(_p(w), _p(R))
    return R.reshape(3, 3)


def o_log(R):
    R = np.ascontiguousarray(R, dtype=np.float64).ravel()
    w = np.zeros(3)
    L.orc_log_so3(_p(R), _p(w))
    return w


def test_exp_matches_matrix_exponential():
    # helpers.h:51-57: skew(axang).exp() — Eigen's Pade matrix exponential == scipy.linalg.expm up to rounding
    for _ in range(200):
        w = rng.normal(0, 1.0, 3) * rng.choice([1e-4, 1e-2, 0.3, 2.0])
        np.testing.assert_allclose(o_exp(w), expm(skew(w)), atol=1e-13, rtol=0)  # Pade scaling-and-squaring carries ~1e-14 itself
        np.testing.assert_allclose(o_exp(w), Rot.from_rotvec(w).as_matrix(), atol=2e-15, rtol=0)


def test_exp_identity_below_epsilon_rot():
    # helpers.h:53-54
    w = np.array([6e-6, 0, 0])
    assert (o_exp(w) == np.eye(3)).all()
    w = np.array([2e-5, 0, 0])
    assert not (o_exp(w) == np.eye(3)).all()


def test_log_matches_matrix_logarithm():
    # helpers.h:59-65: rotm.log() — principal logarithm
    for _ in range(200):
        w = rng.normal(0, 1.0, 3)
        w = w / np.linalg.norm(w) * rng.uniform(1e-6, 3.0)
        R = expm(skew(w))
        lg = np.real(logm(R))
        ref = np.array([lg[2, 1], lg[0, 2], lg[1, 0]])
        np.testing.assert_allclose(o_log(R), ref, atol=5e-12, rtol=0)
        np.testing.assert_allclose(o_log(R), w, atol=5e-12, rtol=0)


def test_log_near_pi_and_zero():
    for ang in (np.pi - 1e-9, np.pi - 1e-4, 3.1):
        for ax in (np.array([1.0, 0, 0]), np.array([0.6, -0.48, 0.64]), np.array([-0.1, 0.2, -0.97])):
            ax = ax / np.linalg.norm(ax)
            R = Rot.from_rotvec(ax * ang).as_matrix()
            w = o_log(R)
            np.testing.assert_allclose(Rot.from_rotvec(w).as_matrix(), R, atol=1e-8)
    assert np.allclose(o_log(np.eye(3)), 0)
    w = np.array([1e-9, -2e-9, 3e-9])
    np.testing.assert_allclose(o_log(expm(skew(w))), w, atol=1e-18, rtol=1e-6)


def test_slerp_matches_scipy():
    # helpers.h:24-37 (Eigen Quaterniond::slerp through AngleAxisd)
    for _ in range(200):
        a = rng.normal(0, 0.5, 3)
        b = a + rng.normal(0, 0.2, 3)
        t = rng.uniform(0, 1)
        o = np.zeros(3)
        L.orc_slerp(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), t, _p(o))
        ref = Slerp([0, 1], Rot.from_rotvec([a, b]))([t]).as_rotvec()[0]
        np.testing.assert_allclose(o, ref, atol=1e-13)
    # identical rotations and the zero rotation
    a = np.array([0.1, 0.2, 0.3])
    o = np.zeros(3)
    L.orc_slerp(_p(a), _p(a.copy()), 0.37, _p(o))
    np.testing.assert_allclose(o, a, atol=1e-15)
    z = np.zeros(3)
    L.orc_slerp(_p(z), _p(z.copy()), 0.5, _p(o))
    assert np.allclose(o, 0)


def fh_weights_bruteforce(x, d):
    # Floater & Hormann 2007, eq. (18): w_k = sum_{i in J_k} (-1)^i prod_{j=i..i+d, j!=k} 1/(x_k - x_j)
    n = len(x)
    w = np.zeros(n)
    for k in range(n):
        for i in range(max(0, k - d), min(k, n - 1 - d) + 1):
            prod = 1.0
            for j in range(i, i + d + 1):
                if j != k:
                    prod *= 1.0 / (x[k] - x[j])
            w[k] += (-1.0) ** i * prod
    return w


@pytest.mark.parametrize("n", [3, 4, 6, 20, 40])
def test_barycentric_rational_weights_and_eval(n):
    # ContinuousTrajectory.h:214 boost::math::barycentric_rational<double>(x, y, n, 2)
    x = np.sort(rng.uniform(0, 2, n))
    x[0] = 0.0
    w = np.zeros(n)
    L.orc_fh_weights(_p(x), n, 2, _p(w))
    np.testing.assert_allclose(w, fh_weights_bruteforce(x, 2), rtol=1e-12)
    # interpolates the nodes exactly and reproduces polynomials of degree <= d
    y = 0.3 - 1.7 * x + 0.9 * x**2
    for i in range(n):
        assert L.orc_fh_eval(_p(x), _p(y), _p(w), n, float(x[i])) == y[i]
    for t in rng.uniform(0, x[-1], 50):
        assert abs(L.orc_fh_eval(_p(x), _p(y), _p(w), n, float(t)) - (0.3 - 1.7 * t + 0.9 * t * t)) < 1e-10
    # linear in y
    y2 = rng.normal(size=n)
    t = float(rng.uniform(0, x[-1]))
    a = L.orc_fh_eval(_p(x), _p(y), _p(w), n, t)
    b = L.orc_fh_eval(_p(x), _p(y2), _p(w), n, t)
    c = L.orc_fh_eval(_p(x), _p(np.ascontiguousarray(y + 2 * y2)), _p(w), n, t)
    assert abs(c - (a + 2 * b)) < 1e-12


@pytest.mark.parametrize("n", [3, 4, 7, 20, 40])
def test_barycentric_rational_matches_scipy_floater_hormann(n):
    """Third-party pin of the Boost restatement: scipy.interpolate.FloaterHormannInterpolator (an independent implementation
    of the same Floater-Hormann family, blending degree d = 2 like ContinuousTrajectory.h:214) must give the same interpolant
    - weights up to the common scale factor of the barycentric form, values to rounding - on the control-pose time grids
    the trajectory model uses (LinSpaced stamps, ContinuousTrajectory.h:332) and on irregular grids."""
    from scipy.interpolate import FloaterHormannInterpolator

    for x in (np.linspace(0.0, 0.1 * (n - 1), n), np.concatenate([[0.0], np.sort(rng.uniform(0.01, 2.0, n - 1))])):
        y = np.stack([np.sin(3 * x) + 0.2 * x, np.cos(2 * x), 0.05 * x**3], axis=1)
        w = np.zeros(n)
        L.orc_fh_weights(_p(x), n, 2, _p(w))
        fh = FloaterHormannInterpolator(x, y, d=2)
        ws = np.asarray(fh.weights, dtype=np.float64).ravel()
        k = int(np.argmax(np.abs(ws)))
        np.testing.assert_allclose(w / w[k], ws / ws[k], rtol=1e-10, atol=1e-13)
        ts = rng.uniform(x[0], x[-1], 40)
        ref = np.asarray(fh(ts))
        for a in range(3):
            ya = np.ascontiguousarray(y[:, a])
            got = np.array([L.orc_fh_eval(_p(x), _p(ya), _p(w), n, float(t)) for t in ts])
            np.testing.assert_allclose(got, ref[:, a], rtol=1e-10, atol=1e-12)


def limit_covariance_numpy(pts):
    """Independent float32 restatement of Gaussians.h:146-154,181-201 with LAPACK's general real eigensolver
    (sgeev — the same class of algorithm as Eigen::EigenSolver<Matrix3f>)."""
    X = pts.astype(np.float32)
    c = X - X.mean(axis=0, dtype=np.float32)
    cov = (c.T @ c) / np.float32(len(X) - 1)
    lam, V = np.linalg.eig(cov.astype(np.float32))
    lam = np.maximum(np.real(lam).astype(np.float32), np.float32(1e-4))
    V = np.real(V).astype(np.float32)
    cov2 = (V @ np.diag(lam) @ np.linalg.inv(V)).astype(np.float32)
    return np.linalg.inv(cov2).astype(np.float32)


def test_gaussian_information_matrix_vs_numpy_float32():
    for trial in range(100):
        n = int(rng.integers(6, 200))
        # planar / linear / blob neighbourhoods incl. clamped directions
        scale = np.array([rng.uniform(0.05, 0.5), rng.uniform(0.05, 0.5), rng.choice([0.001, 0.01, 0.2])])
        R = Rot.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        pts = (rng.normal(size=(n, 3)) * scale) @ R.T + rng.uniform(-20, 20, 3)
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        info = np.zeros(9, dtype=np.float32)
        L.orc_gaussian_info(_p(pts), n, _p(info))
        info = info.reshape(3, 3)
        ref = limit_covariance_numpy(pts)
        cond = np.linalg.cond(ref.astype(np.float64))
        # float32 eigen/inverse noise scales with the condition number (<= ~5e3 with the 1e-4 floor at these scales)
        tol = 4e-6 * cond + 1e-4
        assert np.linalg.norm(info - ref) / np.linalg.norm(ref) < tol, (trial, cond)
        assert np.allclose(info, info.T, rtol=1e-3, atol=1e-3 * np.abs(info).max())
        assert np.all(np.linalg.eigvalsh(0.5 * (info + info.T).astype(np.float64)) > 0)


def test_gaussian_eigenvalue_clamp():
    # points exactly on a plane z = const: smallest covariance eigenvalue 0 -> clamped to 1e-4 -> information 1e4 along z
    n = 50
    pts = np.zeros((n, 3), dtype=np.float32)
    pts[:, 0] = rng.uniform(-1, 1, n)
    pts[:, 1] = rng.uniform(-1, 1, n)
    pts[:, 2] = 2.0
    info = np.zeros(9, dtype=np.float32)
    L.orc_gaussian_info(_p(pts), n, _p(info))
    assert abs(info[8] - 1e4) / 1e4 < 1e-3


def test_pose_chain_roundtrip():
    # ConsecutivePoses.h:26-67
    n = 12
    relO = np.ascontiguousarray(rng.normal(0, 0.2, (n, 3)))
    relT = np.ascontiguousarray(rng.normal(0, 1.0, (n, 3)))
    gO, gT = np.zeros((n, 3)), np.zeros((n, 3))
    L.orc_relative2global(n, _p(relO), _p(relT), _p(gO), _p(gT))
    # independent chain with scipy
    R = np.eye(3)
    T = np.zeros(3)
    for k in range(n):
        T = T + R @ relT[k]
        np.testing.assert_allclose(gT[k], T, atol=1e-13)
        R = R @ Rot.from_rotvec(relO[k]).as_matrix()
        np.testing.assert_allclose(Rot.from_rotvec(gO[k]).as_matrix(), R, atol=1e-13)
    rO, rT = np.zeros((n, 3)), np.zeros((n, 3))
    L.orc_global2relative(n, _p(gO), _p(gT), _p(rO), _p(rT))
    np.testing.assert_allclose(rO[1:], relO[1:], atol=1e-12)
    np.testing.assert_allclose(rT[1:], relT[1:], atol=1e-12)
    np.testing.assert_allclose(rT[0], gT[0])


def test_lu_inverse():
    for n in (1, 5, 18, 114):
        A = rng.normal(size=(n, n))
        A = A.T @ A + 1e-3 * np.eye(n)
        inv = np.zeros((n, n))
        L.orc_lu_inverse(_p(np.ascontiguousarray(A)), n, _p(inv))
        np.testing.assert_allclose(inv @ A, np.eye(n), atol=1e-7)
        np.testing.assert_allclose(inv, np.linalg.inv(A), rtol=1e-6, atol=1e-9)


def lattice(xyzw, res):
    N = len(xyzw)
    ck = np.zeros((N, 3), dtype=np.int32)
    fin = np.zeros(N, dtype=np.uint8)
    lo = np.zeros(3, dtype=np.int64)
    depth = C.c_int32()
    mm = C.c_int64()
    L.orc_lattice(_p(xyzw), N, np.float32(res), _p(ck), _p(fin), _p(lo), C.byref(depth), C.byref(mm))
    return ck, fin, lo, depth.value, mm.value


def test_pcl_lattice_against_bruteforce_octree():
    """PCL semantics restated twice: the oracle's incremental bounding-box growth versus a closed-form lattice
    floor((p - (p0 - res)) / res) and an explicit replay of the root growth in integer key space."""
    for trial in range(20):
        N = 3000
        res = np.float32(rng.choice([0.6, 1.5, 0.25]))
        pts = np.ones((N, 4), dtype=np.float32)
        pts[:, :3] = rng.normal(0, rng.choice([2.0, 15.0]), (N, 3)).astype(np.float32) + rng.uniform(-30, 30, 3).astype(np.float32)
        if trial % 3 == 0:
            pts[5, 0] = np.nan  # PCL skips non-finite points
            pts[77, 2] = np.inf
        ck, fin, lo, depth, mm = lattice(pts, res)
        assert mm == 0
        ok = np.isfinite(pts[:, :3]).all(axis=1)
        assert (fin.astype(bool) == ok).all()
        r = float(res)
        p0 = pts[ok][0, :3].astype(np.float64)
        ref = np.floor((pts[ok, :3].astype(np.float64) - (p0 - r)) / r).astype(np.int64)
        assert (ck[ok] == ref).all()
        # replay root growth in key space (adoptBoundingBoxToPoint)
        l = np.zeros(3, dtype=np.int64)
        d = 1
        for k in ref:
            while True:
                up = k >= l + (1 << d)
                low = k < l
                if not (up.any() or low.any()):
                    break
                l = np.where(~up, l - (1 << d), l)
                d += 1
        assert (l == lo).all() and d == depth
        assert ((ref - lo) >= 0).all() and ((ref - lo) < (1 << depth)).all()


def test_first_point_cell_and_anchor():
    pts = np.ones((4, 4), dtype=np.float32)
    pts[:, :3] = [[1.0, 2.0, 3.0], [1.29, 2.0, 3.0], [1.31, 2.0, 3.0], [0.69, 2.0, 3.0]]
    ck, _, _, _, _ = lattice(pts, 0.6)
    # anchor min = p0 - res: p0 sits exactly in the middle of cell (1,1,1)... key = floor(res/res) = 1
    assert (ck[0] == [1, 1, 1]).all()
    assert ck[1, 0] == 1 and ck[2, 0] == 1 and ck[3, 0] == 0
