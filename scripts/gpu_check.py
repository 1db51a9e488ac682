"""Stage-by-stage comparison of the CUDA path with the CPU oracle (development aid; the pytest suite is the judge)."""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_binding as ob  # noqa: E402
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimizer, DmsaOptimSettings, synth  # noqa: E402

ST = dict(num_iter=5, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30)


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


def stage(name):
    def deco(fn):
        def run(*a, **k):
            t = time.time()
            try:
                r = fn(*a, **k)
                print(f"[{name}] ok ({time.time() - t:.2f}s)", flush=True)
                return r
            except Exception:
                print(f"[{name}] FAILED\n{traceback.format_exc()}", flush=True)
                return None
        return run
    return deco


def check(cfg, full=True):
    print(f"===== {cfg} =====", flush=True)
    win = synth.make_config(cfg)
    st = dict(ST)
    if cfg == "tiny":
        st.update(min_num_points_per_set=6, min_num_gaussians=10)
    s = DmsaOptimSettings(**st)
    so = ob.settings(**st)
    traj = ContinuousTrajectory.from_window(win)
    om = ob.OracleModel.from_window(win)
    om.set_threads(8)
    n_scan = sum(len(x) for x in win["scans"])

    @stage("tform ids")
    def _():
        ids = traj.tformIdPerPoint(n_scan)
        print("   tform id mismatches:", int((ids != om.tid).sum()), "of", n_scan)
        tim = traj.timing()
        print("   n_total", tim["n_total"], om.timing["n_total"], "stamps maxdiff", np.abs(tim["stamps"] - om.timing["stamps"]).max(),
              "trajTime maxdiff", np.abs(tim["traj_time"] - om.timing["traj_time"]).max())
    _()
    traj.centralize()
    om.centralize()

    @stage("global points")
    def _():
        traj.updateGlobalPoints()
        om.update_global_points()
        Mg = traj.denseTforms()
        Mo, _, _ = om.dense_tforms()
        print("   dense tforms: bitwise mismatches", int((Mg.view(np.uint32) != Mo.view(np.uint32)).sum()), "of", Mg.size, "maxabs", np.abs(Mg - Mo).max())
        wg = traj.globalPoints()
        wo = om.world_points()
        print("   world points: bitwise mismatches", int((wg.view(np.uint32) != wo.view(np.uint32)).sum()), "of", wg.size, "maxabs", np.abs(wg - wo).max())
    _()

    @stage("sets")
    def _():
        G, M = traj.buildSets(s)
        Go = om.build_sets(so)
        so_ = om.sets()
        sg = traj.getSets()
        print("   G", G, Go, "M", M, so_["M"], "oracle lattice mismatches", so_["lattice_mismatch"])
        if G == Go and M == so_["M"]:
            print("   offsets equal", bool((sg["offs"] == so_["offs"]).all()), "members equal", bool((sg["members"] == so_["members"]).all()),
                  "keys equal", bool((sg["key"] == so_["key"]).all()), "levels equal", bool((sg["level"] == so_["level"]).all()))
            print("   info rel", rel(sg["info"], so_["info"]), "bitwise mismatches", int((sg["info"].view(np.uint32) != so_["info"].view(np.uint32)).sum()),
                  "w rel", rel(sg["w"], so_["w"]), "w bitwise mismatches", int((sg["w"].view(np.uint32) != so_["w"].view(np.uint32)).sum()))
        for lvl in (0, 1):
            keys, lo, depth = traj.voxelKeys(lvl)
            print(f"   level {lvl}: root lo {lo} depth {depth}")
    _()

    @stage("cost")
    def _():
        p = traj.getPoseParameters()
        po = om.get_params()
        print("   params maxdiff", np.abs(p - po).max())
        eg = traj.evalCost(p[None, :])[0]
        for mode in (2, 0, 1):
            om.set_mode(mode)
            eo = om.cost(po)
            print(f"   e vs oracle mode {mode}: rel {rel(eg, eo):.3e} maxabs {np.abs(eg - eo).max():.3e}")
        om.set_mode(2)
    _()

    @stage("jacobian")
    def _():
        t = time.time()
        cj = traj.costJacobian(with_rows=True)
        print(f"   gpu cost_jacobian wall {time.time() - t:.3f}s")
        for mode in (2, 0, 1):
            om.set_mode(mode)
            e0, J = om.jacobian()
            H = J.T @ J
            g = J.T @ e0
            print(f"   mode {mode}: rel e0 {rel(cj['e0'], e0):.3e} J {rel(cj['J'], J):.3e} H {rel(cj['H'], H):.3e} g {rel(cj['g'], g):.3e}")
        om.set_mode(2)
    if full:
        _()

    @stage("iteration")
    def _():
        p0 = om.get_params()
        for it in range(3):
            d = traj.iteration(s)
            r = om.iteration(so)
            tr = om.last_trace()
            print(f"   it{it}: stop {d['stop']} / {r}  G {d['num_gaussians']}  err0 {d['error0']:.6f} vs {(tr['e0'] ** 2).sum():.6f}  best {d['best_step']} vs {tr['best_k']}"
                  f"  step rel {rel(d['step'], tr['step']):.3e}  ls rel {rel(d['ls_cost'], tr['ls_cost']):.3e}  params rel {rel(traj.getPoseParameters(), om.get_params()):.3e}")
    _()

    @stage("optimize")
    def _():
        t2 = ContinuousTrajectory.from_window(win)
        o2 = ob.OracleModel.from_window(win)
        o2.set_threads(8)
        o2.set_mode(2)
        t = time.time()
        rep = DmsaOptimizer().optimizeSet(t2, s)
        tg = time.time() - t
        t = time.time()
        it, reason = o2.optimize(so)
        to = time.time() - t
        pg, po = t2.getPoses(), o2.get_poses()
        print(f"   gpu {rep['iterations']} it ({rep['stop']}) {tg:.3f}s | oracle {it} it (reason {reason}) {to:.3f}s")
        print("   rel_orient rel", rel(pg["rel_orient"], po["rel_orient"]), "rel_transl rel", rel(pg["rel_transl"], po["rel_transl"]),
              "glob_transl rel", rel(pg["glob_transl"], po["glob_transl"]))
        wg, wo = t2.globalPoints(), o2.world_points()
        print("   final world points maxabs", np.abs(wg - wo).max())
    if full:
        _()
    return traj, s


def timing(cfg, iters=5):
    print(f"===== timing {cfg} =====", flush=True)
    win = synth.make_config(cfg)
    s = DmsaOptimSettings(**ST)
    traj = ContinuousTrajectory.from_window(win)
    traj.centralize()
    for i in range(iters + 2):
        l0 = traj.ctx.launch_count
        t = time.time()
        d = traj.iteration(s)
        traj.ctx.synchronize()
        dt = time.time() - t
        print(f"   iteration {i}: {dt * 1e3:.2f} ms  G {d['num_gaussians']} stop {d['stop']} launches {traj.ctx.launch_count - l0}", flush=True)


if __name__ == "__main__":
    cfgs = sys.argv[1:] or ["tiny", "cfg1"]
    for c in cfgs:
        if c.startswith("time:"):
            timing(c[5:])
        else:
            check(c, full=not c.endswith("-"))
