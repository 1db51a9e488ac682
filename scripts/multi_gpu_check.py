"""Row-sharded keyframe pass across ranks: every rank stages the SAME submap, owns the sets g % world == rank, and
dmsa_b200_iteration all-reduces [H | g | e0^T e0] and the 9 line-search costs with NCCL inside the library.  Rank 0 also runs
the unsharded context on its GPU and compares.  Launch: torchrun --nproc-per-node N scripts/multi_gpu_check.py [n_kf n_pts]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmsa_lidar_slam_b200 import DmsaOptimSettings, MapManagement, synth  # noqa: E402
from dmsa_lidar_slam_b200.api import comm_unique_id  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")  # rendezvous only: the data path is the library's own NCCL communicator
n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
sm = synth.make_keyframe_submap(n_keyframes=n_kf, n_points=n_pts, seed=9)
st = dict(num_iter=3, step_length_optim=0.2, max_step=0.01, min_num_points_per_set=6, min_num_gaussians=10, gauss_split=1, epsilon=1e-7)
s = DmsaOptimSettings(**st)
ids = [comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
kf = MapManagement.from_submap(sm, device=local)
kf.commInit(ids[0], rank, world)
res = [kf.iteration(s) for _ in range(3)]
p = kf.getPoseParameters()
allp = [None] * world
dist.all_gather_object(allp, p.tobytes())
out = dict(world=world, collectives=int(kf.collective_count))
if rank == 0:
    same = all(a == allp[0] for a in allp)  # every rank took the same steps
    ref = MapManagement.from_submap(sm, device=local)
    rr = [ref.iteration(s) for _ in range(3)]
    pr = ref.getPoseParameters()
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
    out.update(ranks_identical=bool(same), G=[d["num_gaussians"] for d in res], G_ref=[d["num_gaussians"] for d in rr],
               best=[d["best_step"] for d in res], best_ref=[d["best_step"] for d in rr],
               err0_rel=[abs(a["error0"] - b["error0"]) / b["error0"] for a, b in zip(res, rr)],
               step_rel=[rel(a["step"], b["step"]) for a, b in zip(res, rr)], ls_rel=[rel(a["ls_cost"], b["ls_cost"]) for a, b in zip(res, rr)],
               params_rel=rel(p, pr))
    ok = same and out["G"] == out["G_ref"] and out["best"] == out["best_ref"] and max(out["err0_rel"]) < 1e-12 and max(out["ls_rel"]) < 1e-10 \
        and out["params_rel"] < 1e-7 and out["collectives"] == 6
    out["ok"] = bool(ok)
kf.commDestroy()
del kf
# ---- bundle mode (BASELINE config 4 shape, small): bundles round-robin over the ranks vs all bundles on one GPU ----
from dmsa_lidar_slam_b200.distributed import KeyframeBundleOptimizer  # noqa: E402

sm2 = synth.make_keyframe_submap(n_keyframes=12, n_points=n_pts, seed=6)
opt = KeyframeBundleOptimizer(sm2, s, 5, 3, rank, world, local, None)
rb = [opt.iteration() for _ in range(3)]
allp = [None] * world
dist.all_gather_object(allp, opt.p.tobytes())
if rank == 0:
    one = KeyframeBundleOptimizer(sm2, s, 5, 3, 0, 1, local, None)
    r1 = [one.iteration() for _ in range(3)]
    out["bundles"] = dict(n_bundles=len(opt.ranges), ranks_identical=bool(all(a == allp[0] for a in allp)),
                          err0_rel=[abs(a["error0"] - b["error0"]) / b["error0"] for a, b in zip(rb, r1)],
                          best=[a["best_step"] for a in rb], best_1gpu=[b["best_step"] for b in r1], params_rel=rel(opt.p, one.p),
                          collectives=opt.collective_count())
    okb = out["bundles"]["ranks_identical"] and max(out["bundles"]["err0_rel"]) < 1e-9 and out["bundles"]["best"] == out["bundles"]["best_1gpu"] \
        and out["bundles"]["params_rel"] < 1e-7
    out["ok"] = bool(out["ok"] and okb)
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
if rank == 0 and not out["ok"]:
    sys.exit(1)
