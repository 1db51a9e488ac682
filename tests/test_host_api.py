"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol the header declares, the
settings/report/point structs have the reference's layout, the synthetic generator is deterministic."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from dmsa_lidar_slam_b200 import api, build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmsa_b200.h")


def test_library_builds_and_exports_every_declared_symbol():
    build.build_library()
    L = api.load_library()
    src = open(HEADER).read()
    declared = sorted(set(re.findall(r"\b(dmsa_b200_[a-z0-9_]+)\s*\(", src)))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/dmsa_b200.h but not exported"
    assert sorted(api.EXPORTED_SYMBOLS) == declared
    assert L.dmsa_b200_version() >= 100


def test_struct_layouts_match_the_header():
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "dmsa_b200.h"
int main(void){
 printf("%zu %zu %zu %zu ", sizeof(dmsa_b200_settings), sizeof(dmsa_b200_point_stamp_id), sizeof(dmsa_b200_point_normal), sizeof(dmsa_b200_report));
 printf("%zu %zu %zu %zu ", offsetof(dmsa_b200_settings, epsilon), offsetof(dmsa_b200_settings, max_step), offsetof(dmsa_b200_settings, lambda_diag), offsetof(dmsa_b200_settings, use_centralization));
 printf("%zu %zu %zu\n", offsetof(dmsa_b200_point_stamp_id, stamp), offsetof(dmsa_b200_point_stamp_id, id), offsetof(dmsa_b200_report, error0));
 return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    v = list(map(int, out))
    S, R = api.DmsaOptimSettings, api.Report
    assert v[0] == C.sizeof(S) and v[1] == synth.POINT_STAMP_ID.itemsize == 32 and v[2] == synth.POINT_NORMAL.itemsize == 48 and v[3] == C.sizeof(R)
    assert v[4:8] == [S.epsilon.offset, S.max_step.offset, S.lambda_diag.offset, S.use_centralization.offset]
    assert v[8] == synth.POINT_STAMP_ID.fields["stamp"][1] == 16 and v[9] == synth.POINT_STAMP_ID.fields["id"][1] == 24
    assert v[10] == R.error0.offset


def test_settings_defaults_are_the_references():
    s = api.DmsaOptimSettings()  # DmsaOptimizer.h:27-38
    assert (s.num_iter, s.epsilon, s.use_analytic_jacobi, s.step_length_optim, s.max_step) == (15, 1e-5, 0, 0.05, 0.01)
    assert (s.gauss_split, s.grid_size_1_factor, s.grid_size_2_factor, s.min_num_points_per_set, s.min_num_gaussians) == (0, 2.0, 5.0, 6, 30)
    assert abs(s.lambda_diag - 1e-5) < 1e-12 and s.use_centralization == 1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.DmsaError, match="no CPU fallback"):
        api.ContinuousTrajectory()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "dmsa_lidar_slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_binding" not in txt and "libdmsa_oracle" not in txt and "/oracle/" not in txt, f


def test_synth_is_deterministic_and_shaped():
    a, b = synth.make_config("tiny"), synth.make_config("tiny")
    assert all((x == y).all() for x, y in zip(a["scans"], b["scans"])) and (a["static"] == b["static"]).all()
    assert (a["rel_orient"] == b["rel_orient"]).all()
    w = synth.make_config("cfg1")
    assert sum(len(s) for s in w["scans"]) == 20000 and len(w["static"]) == 5000 and w["n_poses"] == 4
    pts = np.concatenate(w["scans"])
    assert (pts["w"] == 1).all() and pts["id"].min() == 0 and pts["id"].max() == 19
    r = np.sqrt(pts["x"] ** 2 + pts["y"] ** 2 + pts["z"] ** 2)
    assert r.min() > 0.5 and r.max() < 51.0
    assert (w["static"]["isStatic"] == 1).all() and (w["static"]["stamp"] == -1000.0).all()
    kf = synth.make_keyframe_submap(3, 500, seed=1)
    nn = np.stack([kf["clouds"][0]["nx"], kf["clouds"][0]["ny"], kf["clouds"][0]["nz"]], 1)
    np.testing.assert_allclose(np.linalg.norm(nn, axis=1), 1.0, atol=1e-5)


def test_window_timing_matches_reference_formulas():
    import oracle_binding as ob

    t = ob.window_timing(0.0, 0.9999, 20, 1e-3)
    assert t["n_total"] == 1002 and t["traj_time"][0] == 0 and t["traj_time"][-1] == t["horizon"]
    assert t["param_indices"][0] == 0 and t["param_indices"][-1] == 1001
    ids = ob.tform_ids(np.array([0.0, 0.00049, 0.0005, 0.9999, 5.0]), 0.0, t["traj_time"])
    assert list(ids) == [0, 1, 1, 1000, 1001]


def test_lm_solve_variants():
    """Host LM step (DmsaOptimizer.h:107-128): the helper-thread variant is bit-identical to the serial explicit inverse; the
    Cholesky variant agrees to conditioning; clamp and NaN guard behave like the reference."""
    from dmsa_lidar_slam_b200.api import lm_solve

    os.environ["DMSA_B200_SOLVER_THREADS"] = "1"  # exercise the helper-thread path (off by default)
    rng = np.random.default_rng(5)
    for n in (18, 114, 234):
        J = rng.normal(size=(3 * n, n))
        H = J.T @ J
        g = rng.normal(size=n)
        hg = np.concatenate([H.ravel(), g, [1.0]])
        s = api.DmsaOptimSettings(step_length_optim=0.2, max_step=1e9)
        a, nan_a = lm_solve(s, hg, n, 1)
        for _ in range(3):
            b, nan_b = lm_solve(s, hg, n, 2)
            assert (a == b).all() and not nan_a and not nan_b
        c, _ = lm_solve(s, hg, n, 0)
        ref = -0.2 * np.linalg.solve(H + np.eye(n) * float(np.float32(1e-5)), g)
        np.testing.assert_allclose(a, ref, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(c, ref, rtol=1e-8, atol=1e-12)
        s2 = api.DmsaOptimSettings(step_length_optim=0.2, max_step=1e-6)  # infinity-norm clamp, DmsaOptimizer.h:125-128
        d, _ = lm_solve(s2, hg, n, 1)
        assert abs(np.abs(d).max() - 1e-6) < 1e-18 and np.allclose(d / np.abs(d).max(), a / np.abs(a).max())
    hg[3] = np.nan
    _, nan = lm_solve(api.DmsaOptimSettings(), hg, n, 1)
    assert nan


def _textbook_lm_step(hg, n, lam, alpha):
    """The unblocked operation sequence host_solve.cpp documents (and kernels_solve.cuh reproduces), one IEEE double operation per
    numpy call: LU with partial pivoting (first maximum wins), the columns of P * I substituted forward with ascending j and
    backward with descending j and a multiplication by the reciprocal diagonal, the product (-alpha X) g in ascending column order."""
    a = hg[:n * n].reshape(n, n).copy()
    a[np.arange(n), np.arange(n)] += lam
    g = hg[n * n:n * n + n]
    piv = np.arange(n)
    for k in range(n):
        p = k + int(np.argmax(np.abs(a[k:, k])))  # argmax returns the first maximum
        if p != k:
            a[[k, p]] = a[[p, k]]
            piv[[k, p]] = piv[[p, k]]
        f = a[k + 1:, k] / a[k, k]
        a[k + 1:, k] = f
        a[k + 1:, k + 1:] -= f[:, None] * a[k, k + 1:][None, :]  # (a product, then a difference: two roundings, like the C loop)
    x = np.zeros((n, n))
    x[np.arange(n), piv] = 1.0
    for i in range(n):
        for j in range(i):
            x[i] -= a[i, j] * x[j]
    for j in range(n - 1, -1, -1):
        x[j] = x[j] * (1.0 / a[j, j])
        for i in range(j):
            x[i] -= a[i, j] * x[j]
    step = np.zeros(n)
    for b in range(n):
        step = step + (-alpha * x[:, b]) * g[b]
    return step


@pytest.mark.parametrize("n", [1, 7, 16, 17, 40, 75])
def test_host_lm_step_equals_the_textbook_elimination_bit_for_bit(n):
    """host_solve.cpp runs the elimination in panels of 16 pivots and both substitution sweeps row by row from registers; every
    element must still see the unblocked sequence.  Pinned here against a numpy restatement of that sequence (general matrices,
    so rows are exchanged at most steps)."""
    from dmsa_lidar_slam_b200.api import lm_solve

    rng = np.random.default_rng(100 + n)
    for trial in range(3):
        A = rng.normal(size=(n, n)) if trial < 2 else (lambda J: J.T @ J)(rng.normal(size=(3 * n, n)))
        g = rng.normal(size=n)
        hg = np.concatenate([A.ravel(), g, [1.0]])
        s = api.DmsaOptimSettings(step_length_optim=0.2, max_step=1e300)
        got, nan = lm_solve(s, hg, n, 1)
        want = _textbook_lm_step(hg, n, float(np.float32(1e-5)), 0.2)
        assert not nan and np.array_equal(got, want), np.abs(got - want).max()


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours): stdout carries ONE JSON line with the contract's
    keys even when native libraries write banners to file descriptor 1; works under a torchrun-style environment where only
    rank 0 reports."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="2"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"]
    quiet = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_host_solver_pool_is_race_free_under_thread_sanitizer(tmp_path):
    """The LM solve's helper threads are on by default: two host threads (two contexts) arming, solving and disarming at the
    same time must produce the serial result bit for bit, with ThreadSanitizer silent (the function multi-versioning
    attribute is compiled out for this build: ifunc resolvers run before the sanitizer runtime is up)."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = open(os.path.join(ROOT, "dmsa_lidar_slam_b200", "csrc", "host_solve.cpp")).read()
    marker = '#define DMSA_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))'
    assert marker in src
    (tmp_path / "host_solve_noclone.cpp").write_text(src.replace(marker, "#define DMSA_CLONES"))
    exe = str(tmp_path / "solver_pool_tsan")
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-mavx2", "-ffp-contract=off", "-pthread", "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "solver_pool_tsan.cpp"), str(tmp_path / "host_solve_noclone.cpp")]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "tsan" in (b.stderr + b.stdout).lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, DMSA_B200_SOLVER_THREADS="1"))
    assert "mismatches 0" in r.stdout, r.stdout + r.stderr[-2000:]
    assert "ThreadSanitizer" not in r.stderr and r.returncode == 0, r.stderr[-3000:]


def test_no_kernel_reads_memory_ahead_of_its_dependency_wait():
    """Programmatic dependent launch (csrc/pdl.cuh): a kernel is set up while its predecessor still runs and must not touch
    memory before griddepcontrol.wait (SASS: ACQBULK).  ptxas hoists non-coherent loads (LDG...CONSTANT: __ldg, loads through
    const __restrict__ kernel parameters) above that instruction, which once gave wrong set counts on the GPU — so the shipped
    library must hold no such load outside libdevice's sin / cos reduction tables, and no memory instruction ahead of the wait."""
    import shutil

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not installed")
    lib = build.build_library()
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = re.split(r"\n\s*Function : ", sass)[1:]
    assert len(kernels) >= 60
    trig_tables = ("k_pose_chain", "k_dense_table", "k_dense_poses", "k_normals_knn6")  # Payne-Hanek table of sin / cos / atan2
    for k in kernels:
        name = k.split("\n", 1)[0]
        ins = [re.sub(r"/\*.*?\*/", "", l).strip() for l in k.split("\n") if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l)]
        if "k_chol_solve" in name:  # cooperative launch, never a programmatic dependent
            continue
        wait = [i for i, l in enumerate(ins) if "ACQBULK" in l]
        assert wait and any("PREEXIT" in l for l in ins), f"{name}: no griddepcontrol prologue"
        early = [l for l in ins[: wait[0]] if re.search(r"\b(LDG|LD\.|ST\.|STG|ATOM|ATOMG|RED|LDS|STS)\b", l)]
        assert not early, f"{name}: memory instructions ahead of griddepcontrol.wait: {early[:3]}"
        if not any(t in name for t in trig_tables):
            assert ".CONSTANT" not in k, f"{name}: non-coherent load in a kernel that may be launched as a programmatic dependent"
