"""Generates the golden vectors under tests/golden/ from the CPU oracle (oracle/dmsa_oracle.cpp).

The reference ships no tests or fixtures for this path and cannot be built here, so these vectors
pin the ORACLE (regression) — parity with the reference itself stays "unpinned" and is argued line by line in the
oracle's citations plus the scipy/numpy cross-checks of tests/test_oracle_primitives.py.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_binding as ob  # noqa: E402
from dmsa_lidar_slam_b200 import synth  # noqa: E402

CASES = {
    "tiny": dict(num_iter=3, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=6, min_num_gaussians=10),
    "cfg1": dict(num_iter=3, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30),
}


def make(name):
    st = CASES[name]
    win = synth.make_config(name)
    out = {}
    for mode, tag in ((0, "faithful"), (2, "exactmean")):
        m = ob.OracleModel.from_window(win)
        m.set_mode(mode)
        m.centralize()
        m.update_global_points()
        if mode == 0:
            G = m.build_sets(ob.settings(**st))
            s = m.sets()
            out.update(G=G, M=s["M"], offs=s["offs"], members=s["members"], info=s["info"], w=s["w"], level=s["level"], key=s["key"],
                       world_crc=np.array([np.bitwise_xor.reduce(m.world_points().view(np.uint32).ravel())], dtype=np.uint32),
                       dense_tforms=m.dense_tforms()[0])
        r = m.iteration(ob.settings(**st))
        tr = m.last_trace()
        out.update({f"{tag}_status": r, f"{tag}_e0": tr["e0"], f"{tag}_H": tr["H"], f"{tag}_g": tr["g"], f"{tag}_step": tr["step"],
                    f"{tag}_ls": tr["ls_cost"], f"{tag}_best_k": tr["best_k"], f"{tag}_params_after": m.get_params()})
    # full optimizeSet (3 iterations) in exact-mean mode: final poses
    m = ob.OracleModel.from_window(win)
    m.set_mode(2)
    it, reason = m.optimize(ob.settings(**st))
    p = m.get_poses()
    out.update(opt_iters=it, opt_reason=reason, opt_rel_orient=p["rel_orient"], opt_rel_transl=p["rel_transl"])
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "G", out["G"], "M", out["M"], "->", os.path.getsize(os.path.join(HERE, f"{name}.npz")), "bytes")


if __name__ == "__main__":
    for n in CASES:
        make(n)
