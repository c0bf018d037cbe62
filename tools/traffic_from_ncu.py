"""profiles/r01_traffic.json from an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum) of `python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e` (north-star batch):
average DRAM bytes per launch of the GEMM and attention kernels over one forward."""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]; idx = {n: i for i, n in enumerate(hdr)}
data = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr): continue
    data.setdefault((int(r[idx["ID"]]), r[idx["Kernel Name"]]), {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
out = {}
for key, pat in (("gemm", "gemm_bf16_kernel"), ("attention", "attention_fwd3_kernel")):
    sel = [v for (i, n), v in data.items() if pat in n]
    rd = sum(v["dram__bytes_read.sum"] for v in sel); wr = sum(v["dram__bytes_write.sum"] for v in sel)
    t = sum(v["gpu__time_duration.sum"] for v in sel)
    out[key] = {"dram_bytes_per_launch": (rd + wr) / len(sel), "launches": len(sel), "dram_read_bytes": rd, "dram_write_bytes": wr,
                "time_share_of_forward": t / sum(v["gpu__time_duration.sum"] for v in data.values()),
                "note": f"ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the {len(sel)} launches of one forward at B=4096 "
                        f"({sys.argv[1].split('/')[-1]})"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
