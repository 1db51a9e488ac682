"""Print the launches of the last iteration found in an ncu launch list (csv of gpu__time_duration.sum)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
k = rows[h].index("Kernel Name"); v = rows[h].index("Metric Value"); g = rows[h].index("Grid Size"); b = rows[h].index("Block Size")
data = rows[h + 2:]
idx = [i for i, r in enumerate(data) if "k_make_fd_batch" in r[k]]
st, en = (idx[-2], idx[-1]) if len(idx) >= 2 else (0, len(data))
tot = 0.0
for r in data[st:en]:
    tot += float(r[v])
    print(f"{r[0]:>5} {r[k][:64]:64} {float(r[v]) / 1e3:9.2f} us  {r[g]} {r[b]}")
print("launches", en - st, "sum", tot / 1e3, "us")
