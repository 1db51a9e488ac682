"""Instruction counts of the shipped libdmsa_b200.so per kernel (cuobjdump -sass) -> profiles/r02_sass_summary.txt.
python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmsa_lidar_slam_b200 import build  # noqa: E402

WATCH = ["UBLKCP", "SYNCS", "DMMA", "FMUL2", "FADD2", "FFMA2", "FFMA", "FMUL", "FADD", "F2F.F64.F32", "DADD", "DMUL", "DFMA", "MATCH.ANY", "SHFL", "LDG", "STG", "LDS", "STS",
         "ATOMG", "ATOMS", "BAR.SYNC", "NANOSLEEP", "PREEXIT", "ACQBULK"]
KERNELS = ["k_cost_fused2<32, 24>", "k_cost_sum2<32, 24>", "k_cost_quad2<32, 24>", "k_cost_fused<true, 64, 16>", "k_jtj_dmma", "k_lu128<false>", "k_inv128", "k_step_fin", "k_sort_prepare",
           "k_sort_pass", "k_segment", "k_emit<false>", "k_emit<true>", "k_scan_excl", "k_keys", "k_root", "k_pose_chain<0>", "k_decode_pc2", "k_normals_knn6"]

lib = build.build_library()
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
parts = re.split(r"\n\s*Function : ", sass)[1:]
print("SASS instruction counts of the shipped libdmsa_b200.so (cuobjdump -sass, sm_100a), selected kernels (scripts/sass_summary.py).")
print("Checks: UBLKCP + SYNCS = TMA bulk copy + mbarrier in the cost kernels; DMMA = FP64 tensor-core J^T J; FMUL2 / FADD2 = packed FP32x2;")
print("FFMA2 = 0 and FFMA only inside division / sqrt sequences (no fused multiply-add on the parity-critical path); MATCH.ANY = warp multi-split of the")
print("radix sort; PREEXIT + ACQBULK = griddepcontrol.launch_dependents / .wait (programmatic dependent launch, csrc/pdl.cuh) at the top of every kernel;")
print("no LDG...CONSTANT (non-coherent load) outside libdevice's trigonometric tables (tests/test_host_api.py checks both on every build).\n")
tot = collections.Counter()
nc = 0
for name, body in zip(names, parts):
    ins = [re.sub(r"/\*.*?\*/", "", l).strip() for l in body.split("\n") if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l)]
    ops = collections.Counter()
    for l in ins:
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
        if not m:
            continue
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                ops[w] += 1
                break
    tot.update(ops)
    nc += body.count(".CONSTANT")
    if any(k in name for k in KERNELS):
        print(name)
        print(f"   instructions {len(ins)}: " + ", ".join(f"{k} {v}" for k, v in ops.items()))
print(f"\nwhole library: {len(parts)} kernels; " + ", ".join(f"{k} {tot[k]}" for k in ("UBLKCP", "DMMA", "FMUL2", "FADD2", "FFMA2", "MATCH.ANY", "PREEXIT", "ACQBULK")) + f", LDG.CONSTANT {nc}")
