"""Top SASS lines by stall samples of one kernel of an .ncu-rep (source page, csv)."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hs = [i for i, r in enumerate(rows) if "Source" in r]
h = hs[0]
end = hs[1] - 1 if len(hs) > 1 else len(rows)
hdr = rows[h]
si = hdr.index("# Samples"); so = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stalls = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[h + 1:end] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
print(rows[h - 1][:2], "total samples", tot, "instructions", len(body))
for pos, r in enumerate(body):
    r.append(pos)
for r in sorted(body, key=lambda r: -int(r[si]))[:top]:
    st = sorted(((int(r[i]), hdr[i]) for i in stalls if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(f"{r[-1]:5d} {int(r[si]):6d} {100.0 * int(r[si]) / max(tot, 1):5.1f}%  {r[so].strip()[:70]:70} {st}")
agg = {}
for r in body:
    for i in stalls:
        if r[i].isdigit() and int(r[i]) > 0:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1]))
# cumulative samples by instruction position (10 bins)
n = len(body)
bins = [0] * 10
for r in body:
    bins[min(9, r[-1] * 10 // max(n, 1))] += int(r[si])
print("samples by tenth of the instruction stream:", bins)
