"""Top stalled SASS instructions of an `ncu --page source --csv --print-source sass` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows if r and r[0].startswith("0x")]
tot = sum(int(r[ix["# Samples"]]) for r in body)
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {c: sum(int(r[ix[c]]) for r in body) for c in stall_cols}
print("total samples", tot, "instructions", len(body))
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:n_top]:
    st = sorted(((int(r[ix[c]]), c) for c in stall_cols), reverse=True)[:2]
    print(f"{int(r[ix['# Samples']]):7d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[1].strip()[:64]:64s} {st}")
