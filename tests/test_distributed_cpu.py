"""The multi-rank protocol of dmsa_lidar_slam_b200/distributed.py on CPU: world_size 2, gloo backend.

The per-rank compute stand-in is the CPU oracle (row-sharded exactly like dmsa_b200_set_shard: rank r owns rows
g % world == r); what is under test is the host logic around it: bundle <-> global parameter mapping, the scatter of
bundle systems into the global [H | g | err0] buffer, the two all-reduces, the redundant LM solve and the step choice."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_binding as ob
from dmsa_lidar_slam_b200 import distributed as D
from dmsa_lidar_slam_b200 import synth
from dmsa_lidar_slam_b200.api import DmsaOptimSettings, lm_solve

ST = dict(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=6, min_num_gaussians=10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_bundle_ranges_and_parameter_mapping():
    assert D.bundle_ranges(64, 15, 8) == [(0, 14), (7, 21), (14, 28), (21, 35), (28, 42), (35, 49), (42, 56), (49, 63)]
    assert D.bundle_ranges(5, 15, 8) == [(0, 4)]
    n = 10
    # one bundle over everything == the reference's single submap: identity mapping
    assert (D.bundle_param_index(n, 0, n - 1) == np.arange(6 * (n - 1))).all()
    idx = D.bundle_param_index(n, 3, 6)
    assert len(idx) == 18
    # orientation block of keyframe 4 (bundle-local pose 1), translation block of keyframe 6 (local pose 3)
    assert list(idx[:3]) == [9, 10, 11] and list(idx[9 + 6:9 + 9]) == [27 + 15, 27 + 16, 27 + 17]
    # params <-> relative poses round trip (Poses.h:64-76)
    rng = np.random.default_rng(0)
    ro, rt = rng.normal(size=(3, n)), rng.normal(size=(3, n))
    p = D.rel_to_params(ro, rt)
    ro2, rt2 = D.params_to_rel(p, ro[:, 0], rt[:, 0], n)
    assert (ro2 == ro).all() and (rt2 == rt).all()
    # scatter positions
    P = 6 * (n - 1)
    pos, Pb = D.hg_scatter_index(idx, P)
    Hb = rng.normal(size=(Pb, Pb))
    gb = rng.normal(size=Pb)
    buf = np.zeros(P * P + P + 1)
    np.add.at(buf, pos, np.concatenate([Hb.ravel(), gb, [2.5]]))
    H = buf[:P * P].reshape(P, P)
    assert (H[np.ix_(idx, idx)] == Hb).all() and (buf[P * P:P * P + P][idx] == gb).all() and buf[-1] == 2.5
    assert np.count_nonzero(H) == Pb * Pb


def test_relative2global_matches_oracle():
    rng = np.random.default_rng(1)
    n = 7
    ro, rt = rng.normal(0, 0.2, (3, n)), rng.normal(0, 1.0, (3, n))
    go, gt = D.relative2global(ro, rt)
    gO, gT = np.zeros((n, 3)), np.zeros((n, 3))
    ob.lib().orc_relative2global(n, ob._p(np.ascontiguousarray(ro.T)), ob._p(np.ascontiguousarray(rt.T)), ob._p(gO), ob._p(gT))
    np.testing.assert_allclose(go.T, gO, atol=1e-12)
    np.testing.assert_allclose(gt.T, gT, atol=1e-12)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        win = synth.make_config("tiny")
        m = ob.OracleModel.from_window(win)
        m.set_mode(2)
        m.centralize()
        m.update_global_points()
        so = ob.settings(**ST)
        m.build_sets(so)
        e0, J = m.jacobian()  # every rank holds the replicated points/parameters; it OWNS rows g % world == rank
        own = (np.arange(len(e0)) % world) == rank
        Jr, er = J[own], e0[own]
        P = J.shape[1]
        part = np.concatenate([(Jr.T @ Jr).ravel(), Jr.T @ er, [er @ er]])
        ex = D.Exchange(world)
        hg = ex.all_reduce_sum(torch.from_numpy(part.copy())).numpy()
        s = DmsaOptimSettings(**ST)
        step, nan = lm_solve(s, hg, P)
        assert not nan
        # line search: partial costs over the owned rows, second exchange
        p = m.get_params()
        ls_part = np.array([float((m.cost(p + 0.1 * k * step)[own] ** 2).sum()) for k in range(1, 10)])
        ls = ex.all_reduce_sum(torch.from_numpy(ls_part.copy())).numpy()
        best = D.select_step(float(hg[-1]), ls)
        if rank == 0:
            np.savez(out, hg=hg, step=step, ls=ls, best=best)
        gathered = [torch.zeros(P, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(step.copy()))
        assert all((g.numpy() == step).all() for g in gathered), "ranks must solve to bit-identical steps"
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_iteration_over_gloo(tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    # single-process reference: the oracle's own iteration
    win = synth.make_config("tiny")
    m = ob.OracleModel.from_window(win)
    m.set_mode(2)
    m.centralize()
    so = ob.settings(**ST)
    assert m.iteration(so) == 0
    tr = m.last_trace()
    P = len(tr["g"])
    H = r["hg"][:P * P].reshape(P, P) + np.eye(P) * float(np.float32(1e-5))
    np.testing.assert_allclose(H, tr["H"], rtol=1e-9, atol=1e-9 * np.abs(tr["H"]).max())
    np.testing.assert_allclose(r["hg"][P * P:P * P + P], tr["g"], rtol=1e-9, atol=1e-9 * np.abs(tr["g"]).max())
    np.testing.assert_allclose(r["step"], tr["step"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(r["ls"], tr["ls_cost"], rtol=1e-8)
    assert int(r["best"]) == tr["best_k"]


def test_library_pose_chain_matches_the_scipy_chain():
    """The bundle driver places every bundle's first keyframe with the library's host pose chain (ConsecutivePoses.h:26-43);
    scipy's rotation algebra is the independent check."""
    import numpy as np

    from dmsa_lidar_slam_b200 import distributed as d

    rng = np.random.default_rng(4)
    ro, rt = rng.normal(0, 0.4, (3, 64)), rng.normal(0, 2.0, (3, 64))
    a, b = d.relative2global(ro, rt), d.relative2global_scipy(ro, rt)
    assert np.abs(a[0] - b[0]).max() < 1e-12 and np.abs(a[1] - b[1]).max() < 1e-11
