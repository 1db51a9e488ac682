"""The header-only C++ adapter (DmsaOptimizer-shaped, dmsa_lidar_slam_b200/host/DmsaOptimizerB200.h) over the C-ABI."""
import os
import struct
import subprocess

import numpy as np
import pytest

from dmsa_lidar_slam_b200 import build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_smoke")


def compile_adapter():
    src = os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cpp")
    lib = build.build_library()
    hdr = os.path.join(ROOT, "dmsa_lidar_slam_b200", "host", "DmsaOptimizerB200.h")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(lib)):
        cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cc, "-std=c++17", "-O1", "-o", EXE, src, "-L" + os.path.dirname(lib), "-ldmsa_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def test_adapter_compiles_against_the_c_abi():
    assert os.path.exists(compile_adapter())


@pytest.mark.gpu
def test_adapter_optimizeSet_matches_python_binding(tmp_path):
    from dmsa_lidar_slam_b200 import ContinuousTrajectory, DmsaOptimizer, DmsaOptimSettings

    exe = compile_adapter()
    win = synth.make_config("tiny")
    inp, out = str(tmp_path / "win.bin"), str(tmp_path / "res.bin")
    n = win["n_poses"]
    with open(inp, "wb") as f:
        f.write(struct.pack("<iiiqddd", len(win["scans"]), n, 3, len(win["static"]), win["t_min"], win["t_max"], win["dt_res"]))
        for sc, gs in zip(win["scans"], win["grid_sizes"]):
            f.write(struct.pack("<qf", len(sc), gs))
            f.write(sc.tobytes())
        f.write(win["static"].tobytes())
        f.write(np.ascontiguousarray(win["rel_orient"].T).tobytes())
        f.write(np.ascontiguousarray(win["rel_transl"].T).tobytes())
    subprocess.check_call([exe, inp, out])
    raw = open(out, "rb").read()
    iters, stop = struct.unpack("<ii", raw[:8])
    arr = np.frombuffer(raw[8:], dtype=np.float64).reshape(3, n, 3)
    traj = ContinuousTrajectory.from_window(win)
    rep = DmsaOptimizer().optimizeSet(traj, DmsaOptimSettings(num_iter=3, step_length_optim=0.2, max_step=0.3, min_num_points_per_set=6, min_num_gaussians=10))
    p = traj.getPoses()
    assert (iters, stop) == (rep["iterations"], rep["stop_reason"])
    assert (arr[0].T == p["rel_orient"]).all() and (arr[1].T == p["rel_transl"]).all() and (arr[2].T == p["glob_transl"]).all()
