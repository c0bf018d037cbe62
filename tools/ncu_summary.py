"""Text summary of an .ncu-rep (raw page): the metrics DESIGN.md / bench.py quote, per captured launch,
followed by the top stalled SASS instructions of the first launch.
    python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep "<title>" > profiles/<name>.txt"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
print("#", title)
print("# ncu --set full --clock-control none --import-source on;", rep.split("/")[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("---")
    for k in KEYS:
        if k in d and d[k] != "":
            unit = rows[1][hdr.index(k)]
            print(f"{k} = {d[k]} {unit}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h2 = next(r for r in srows if r and r[0] == "Address")
ix = {n: i for i, n in enumerate(h2)}
body = []
for r in srows[srows.index(h2) + 1:]:
    if r and r[0] == "Kernel Name": break
    if r and r[0].startswith("0x"): body.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in body)
stall_cols = [n for n in h2 if n.startswith("stall_") and "Not Issued" not in n]
agg = sorted(((sum(int(r[ix[c]]) for r in body), c) for c in stall_cols), reverse=True)
print("--- warp-state samples of the first launch:", tot, "over", len(body), "SASS instructions")
print("by reason:", ", ".join(f"{c.replace('stall_', '')} {100 * v / tot:.1f}%" for v, c in agg[:9]))
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
    st = sorted(((int(r[ix[c]]), c.replace("stall_", "")) for c in stall_cols), reverse=True)[0]
    print(f"{100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[1].strip()[:72]:72s} {st[1]}")
