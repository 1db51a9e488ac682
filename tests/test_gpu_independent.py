"""The CUDA path against the INDEPENDENT numpy / scipy model of tests/test_independent_model.py — not through the oracle, so a
mistake the oracle and the kernels might share (their SO(3) code is a twin) cannot hide here.  `-m gpu`.

Tolerances: membership identical; everything else up to float noise (the numpy model is float64 but for the point transform,
the kernels are the reference's float arithmetic): e0 1e-5, information matrices 1e-6 median, H and g inside the north star's
1e-4, J 1e-3 (a forward difference over h = 3.5e-4 amplifies float noise ~3 000 x)."""
import numpy as np
import pytest

import test_independent_model as im
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings, MapManagement, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["tiny", "cfg1"])
def test_sliding_window_pass_against_the_numpy_model(name):
    win = synth.make_config(name)
    st = dict(im.SETTINGS, min_num_points_per_set=6 if name == "tiny" else 10)
    traj = ContinuousTrajectory.from_window(win)
    traj.centralize()
    traj.updateGlobalPoints()
    G, M = traj.buildSets(DmsaOptimSettings(**st))
    # the same centralisation on the model (ContinuousTrajectory.h:75-87): pose 0 to the origin, static points shifted in float
    mdl = im.NumpyTrajectoryModel(win)
    origin = mdl.rel_transl0[:, 0].copy()
    mdl.rel_transl0[:, 0] = 0.0
    mdl.static = (mdl.static - origin.astype(np.float32)).astype(np.float32)
    p0 = mdl.params()
    assert im.rel(p0, traj.getPoseParameters()) == 0.0
    world, _, _ = mdl.world(p0)
    wg = traj.globalPoints()[:, :3]
    assert np.abs(wg - world).max() <= 8e-6 * max(1.0, np.abs(world).max() / 40.0)
    sets_np = []
    for f in (2.0, 5.0):
        sets_np += im.voxel_sets(wg, mdl.ring, np.float32(f) * np.float32(mdl.min_grid), st["min_num_points_per_set"])  # on the GPU's floats
    sg = traj.getSets()
    assert len(sets_np) == G == sg["G"]
    idx = im.match(sets_np, im.oracle_sets_as_lists(sg))  # KeyError = a set of the CUDA path the lattice restatement does not have
    assert len(set(idx.tolist())) == G
    infos, wt = im.gaussians(wg, sets_np)
    info_np = np.stack([infos[i] for i in idx]).reshape(G, 9)
    err = np.linalg.norm(sg["info"].astype(np.float64) - info_np, axis=1) / np.linalg.norm(info_np, axis=1)
    assert np.median(err) < 1e-6 and err.max() < 1e-3, (np.median(err), err.max())
    assert im.rel(sg["w"], wt[idx]) < 1e-6
    cj = traj.costJacobian(with_rows=True)
    e_np = im.residuals(world, sets_np, infos, wt)[idx]
    assert im.rel(e_np, cj["e0"]) < 5e-5, im.rel(e_np, cj["e0"])
    h = float(np.sqrt(np.float64(np.finfo(np.float32).eps)))
    J = np.stack([(im.residuals(mdl.world(p0 + h * np.eye(len(p0))[k])[0], sets_np, infos, wt)[idx] - e_np) / h for k in range(len(p0))], axis=1)
    assert im.rel(J, cj["J"]) < 1e-3, im.rel(J, cj["J"])
    assert im.rel(J.T @ J, cj["H"]) < 1e-4, im.rel(J.T @ J, cj["H"])
    assert im.rel(J.T @ e_np, cj["g"]) < 1e-4, im.rel(J.T @ e_np, cj["g"])
    print(f"\n[{name}] CUDA vs numpy model: world {np.abs(wg - world).max():.2e}, info median {np.median(err):.1e} max {err.max():.1e}, e0 {im.rel(e_np, cj['e0']):.1e}, "
          f"J {im.rel(J, cj['J']):.1e}, H {im.rel(J.T @ J, cj['H']):.1e}, g {im.rel(J.T @ e_np, cj['g']):.1e}")


def test_keyframe_pass_with_split_sets_against_the_numpy_model():
    sm = synth.make_keyframe_submap(n_keyframes=5, n_points=6000, seed=9)
    st = dict(num_iter=1, step_length_optim=0.2, max_step=0.01, min_num_points_per_set=6, min_num_gaussians=10, gauss_split=1, epsilon=1e-4)
    kf = MapManagement.from_submap(sm)
    kf.updateGlobalPoints()
    G, M = kf.buildSets(DmsaOptimSettings(**st))
    ro, rt = sm["rel_orient"].astype(np.float64), sm["rel_transl"].astype(np.float64)
    world, normals = im.keyframe_world(sm, ro, rt)
    wg, ng = kf.globalPoints(normals=True)
    wg, ng = wg[:, :3], ng[:, :3]
    assert np.abs(wg - world).max() <= 1.6e-5 and np.abs(ng - normals).max() <= 4e-7
    ring = np.concatenate(sm["rings"]).astype(np.int64)
    grid = np.float32(min(sm["grid_sizes"]))
    sets_np = []
    for f in (2.0, 5.0):
        sets_np += im.keyframe_sets(wg, ng, ring, np.float32(f) * grid, st["min_num_points_per_set"], 1)
    sg = kf.getSets()
    assert (sg["sub"] > 0).sum() > 10 and len(sets_np) == G
    idx = im.match(sets_np, im.oracle_sets_as_lists(sg))
    assert len(set(idx.tolist())) == G
    infos, wt = im.gaussians(wg, sets_np)
    cj = kf.costJacobian(with_rows=True)
    e_np = im.residuals(world, sets_np, infos, wt)[idx]
    assert im.rel(e_np, cj["e0"][:G]) < 5e-5, im.rel(e_np, cj["e0"][:G])
    print(f"\n[keyframe, split] CUDA vs numpy model: world {np.abs(wg - world).max():.2e}, normals {np.abs(ng - normals).max():.1e}, G {G}, e0 {im.rel(e_np, cj['e0'][:G]):.1e}")
