"""Instruction mix / hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME` (SASS view)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ishw, ishi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data); ts = sum(int(r[isamp]) for r in data)
print(f"total warp instructions {tot / 1e9:.3f} G, stall samples {ts}, SASS lines {len(data)}")
byop, bys = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ia]); op = m.group(2) if m else "?"
    byop[op] += int(r[ie]); bys[op] += int(r[isamp])
for op, c in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:10s} {c / 1e6:9.1f} M  {100 * c / tot:5.1f} % of instructions   {100 * bys[op] / max(ts, 1):5.1f} % of stall samples")
print(f"shared-memory wavefronts {sum(int(r[ishw]) for r in data) / 1e9:.3f} G (ideal {sum(int(r[ishi]) for r in data) / 1e9:.3f} G)")
