"""How far can the float noise of the reference's information matrices move e0 / J / H / g?   

Gaussians.h:146-154,181-201 computes the covariance, its eigen-decomposition (EigenSolver<Matrix3f>: Hessenberg + shifted QR),
V L V^-1 and the inverse in FLOAT; Eigen is not on this machine, so the oracle takes the exactly-rounded value of that
chain and nobody can say which float the real reference lands on.  This script bounds the consequence with the independent numpy
model (tests/test_independent_model.py): the same sets, once with information matrices from a float64 symmetric
decomposition (what the oracle's values round to) and once from an all-float32 chain whose eigen-decomposition is LAPACK's
general real solver sgeev (numpy.linalg.eig on float32 — the same algorithm family and precision as EigenSolver<Matrix3f>),
and reports the relative change of e0, J, H = J^T J and g = J^T e0.   CPU only:  python scripts/eigen_noise_sensitivity.py [cfg1 cfg2]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_independent_model as im  # noqa: E402
from dmsa_lidar_slam_b200 import synth  # noqa: E402


def gaussians_float32(world, sets):
    infos = []
    for m in sets:
        x = world[m].astype(np.float32)
        c = (x - x.mean(axis=0, dtype=np.float32)).astype(np.float32)
        cov = (c.T @ c / np.float32(len(m) - 1)).astype(np.float32)
        lam, V = np.linalg.eig(cov)  # float32 in, sgeev
        lam, V = np.real(lam).astype(np.float32), np.real(V).astype(np.float32)
        lam = np.maximum(lam, np.float32(1e-4))
        cov2 = (V @ np.diag(lam) @ np.linalg.inv(V).astype(np.float32)).astype(np.float32)
        infos.append(np.linalg.inv(cov2).astype(np.float32).astype(np.float64))
    return infos


out = {}
for name in sys.argv[1:] or ["cfg1"]:
    t0 = time.time()
    win = synth.make_config(name)
    mdl = im.NumpyTrajectoryModel(win)
    p0 = mdl.params()
    world = mdl.world(p0)[0]
    minpts = 6 if name == "tiny" else 10
    sets = []
    for f in (2.0, 5.0):
        sets += im.voxel_sets(world, mdl.ring, np.float32(f) * np.float32(mdl.min_grid), minpts)
    info64, wt = im.gaussians(world, sets)
    info32 = gaussians_float32(world, sets)
    d_info = np.array([np.linalg.norm(a - b) / np.linalg.norm(a) for a, b in zip(info64, info32)])
    h = float(np.sqrt(np.float64(np.finfo(np.float32).eps)))
    worlds = [world] + [mdl.world(p0 + h * np.eye(len(p0))[k])[0] for k in range(len(p0))]  # shared by both variants
    res = {}
    for tag, infos in (("f64", info64), ("f32", info32)):
        E = np.stack([im.residuals(w, sets, infos, wt) for w in worlds], axis=1)
        e0 = E[:, 0]
        J = (E[:, 1:] - e0[:, None]) / h
        res[tag] = dict(e0=e0, J=J, H=J.T @ J, g=J.T @ e0)
    # the same model against the FAITHFUL oracle at this size (sets matched by their members)
    import oracle_binding as ob
    om = ob.OracleModel.from_window(win)
    om.set_threads(os.cpu_count() or 8)
    om.set_mode(0)
    om.update_global_points()
    G = om.build_sets(ob.settings(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=minpts, min_num_gaussians=10))
    idx = im.match(sets, im.oracle_sets_as_lists(om.sets()))
    e_or, J_or = om.jacobian()
    vs_oracle = dict(sets_identical=bool(G == len(sets)), e0=im.rel(res["f64"]["e0"][idx], e_or), J=im.rel(res["f64"]["J"][idx], J_or),
                     H=im.rel(res["f64"]["H"], J_or.T @ J_or), g=im.rel(res["f64"]["g"], J_or.T @ e_or))
    out[name] = dict(G=len(sets), P=len(p0), numpy_model_vs_faithful_oracle=vs_oracle, info_rel_median=float(np.median(d_info)), info_rel_p99=float(np.percentile(d_info, 99)), info_rel_max=float(d_info.max()),
                     **{k: im.rel(res["f32"][k], res["f64"][k]) for k in ("e0", "J", "H", "g")}, seconds=round(time.time() - t0, 1))
    print(name, json.dumps(out[name]), flush=True)
print(json.dumps(dict(what="relative change of e0 / J / H / g when the information matrices come from an all-float32 chain with LAPACK sgeev instead of a float64 "
                           "symmetric decomposition, same sets, independent numpy model (scripts/eigen_noise_sensitivity.py)", **out)))
