"""What would fusing multiply-adds do to H and g?  

Builds a second library with FMA contraction allowed on the cost path (-fmad=true, plain float operators instead of the
__fmul_rn / __fadd_rn intrinsics), evaluates e0 / J / H / g of one BASELINE config with the scalar cost kernels of both
libraries, and writes the relative deviations next to the distance of the shipped library from the oracle's faithful arithmetic."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimSettings, synth
win = synth.make_config(%r)
s = DmsaOptimSettings(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30)
t = ContinuousTrajectory.from_window(win)
t.setPairMode(0)
t.centralize(); t.updateGlobalPoints(); t.buildSets(s)
cj = t.costJacobian(with_rows=True)
np.savez(sys.argv[1], e0=cj["e0"], J=cj["J"], H=cj["H"], g=cj["g"])
"""

if __name__ == "__main__":
    from dmsa_lidar_slam_b200 import build

    fma = build.OUT_FMA
    if not os.path.exists(fma) or any(os.path.getmtime(d) > os.path.getmtime(fma) for d in build.DEPS):
        fma = build.build_fma_variant()
    build.build_library()
    outs = {}
    for tag, lib in (("faithful", None), ("fma", fma)):
        env = dict(os.environ)
        if lib:
            env["DMSA_B200_LIB"] = lib
        path = f"/tmp/fma_dev_{tag}.npz"
        subprocess.check_call([sys.executable, "-c", CHILD % (ROOT, cfg), path], env=env)
        outs[tag] = dict(np.load(path))
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    res = {"config": cfg, "what": "scalar cost kernels, same sets; 'fma' = the same source compiled with -fmad=true and plain float operators (FFMA contraction)",
           "fma_vs_shipped": {k: rel(outs["fma"][k], outs["faithful"][k]) for k in ("e0", "J", "H", "g")}}
    import oracle_binding as ob
    from dmsa_lidar_slam_b200 import synth

    win = synth.make_config(cfg)
    om = ob.OracleModel.from_window(win)
    om.set_threads(os.cpu_count() or 8)
    om.set_mode(0)
    om.centralize(); om.update_global_points()
    om.build_sets(ob.settings(num_iter=1, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=10, min_num_gaussians=30))
    e0, J = om.jacobian()
    H, g = J.T @ J, J.T @ e0
    for tag in ("faithful", "fma"):
        o = outs[tag]
        res[("shipped" if tag == "faithful" else "fma") + "_vs_oracle_faithful"] = {"e0": rel(o["e0"], e0), "J": rel(o["J"], J), "H": rel(o["H"], H), "g": rel(o["g"], g)}
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"fma_deviation_{cfg}.json"), "w"), indent=1)
