"""Text summary of an .ncu-rep (raw page): per launch duration, DRAM bytes, achieved DRAM GB/s, issue / pipe utilisation, occupancy."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "ns"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("smsp__inst_executed.sum", "warp_insts"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak")]
units = rows[1]
for r in rows[2:]:
    print("-" * 100)
    d = {}
    for name, short in want:
        if name in col:
            d[short] = r[col[name]]
    def num(k):
        try:
            return float(d[k].replace(",", ""))
        except Exception:
            return None
    ns = num("ns")
    rd, wr = num("dram_rd"), num("dram_wr")
    # units of dram bytes can be Kbyte / Mbyte in the csv: normalise with the unit row
    def scale(name):
        u = units[col[name]] if name in col else ""
        return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if rd is not None:
        rd *= scale("dram__bytes_read.sum")
    if wr is not None:
        wr *= scale("dram__bytes_write.sum")
    tu = units[col["gpu__time_duration.sum"]]
    if ns is not None and tu in ("us", "usecond"):
        ns *= 1e3
    elif ns is not None and tu in ("ms", "msecond"):
        ns *= 1e6
    print(f"{d.get('kernel', '')[:70]}  grid {d.get('grid')} block {d.get('block')} regs {d.get('regs')}")
    if ns:
        print(f"   duration {ns / 1e3:.2f} us; DRAM read {rd / 1e6:.3f} MB, write {wr / 1e6:.3f} MB -> {(rd + wr) / ns:.1f} GB/s ({d.get('dram_pct_of_peak')} % of peak)")
    print("   " + ", ".join(f"{k} {d[k]}" for k in ("warps_active_pct", "issue_active_pct", "fma_pipe_pct", "xu_pipe_pct", "fp64_pipe_pct", "dmma_pipe_pct", "l1_hit_pct", "l2_hit_pct", "warp_insts") if k in d))
