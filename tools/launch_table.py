"""Prints a per-launch table (time, share, DRAM bytes, tensor-pipe %) from an ncu --csv launch list."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]; idx = {n: i for i, n in enumerate(hdr)}
data = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr): continue
    data.setdefault((int(r[idx["ID"]]), r[idx["Kernel Name"]]), {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
tot = sum(v["gpu__time_duration.sum"] for v in data.values())
agg = collections.defaultdict(float)
for (i, name), v in data.items():
    t = v["gpu__time_duration.sum"]
    short = name.split("(")[0].replace("void ", "").replace("hriemo::", "")[:44]
    agg[short] += t
    print(f"{i:3d} {short:44s} {t/1e3:9.1f}us {100*t/tot:5.1f}%  rd {v.get('dram__bytes_read.sum',0)/1e6:8.1f}MB wr {v.get('dram__bytes_write.sum',0)/1e6:8.1f}MB"
          f"  tensor {v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',0):5.1f}%  cyc {v.get('sm__cycles_elapsed.avg',0)/1e3:8.0f}k")
print(f"total {tot/1e6:.3f} ms")
for k, t in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {k:44s} {t/1e6:8.3f} ms {100*t/tot:5.1f}%")
