"""Instruction mix and stall samples by opcode from `ncu -i X.ncu-rep --page source --csv` (SASS view)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
mix = collections.Counter(); smp = collections.Counter(); tot = 0; ts = 0
for r in rows[2:]:
    if len(r) <= iE: continue
    op = r[iS].split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.split(".")[0] if not o.startswith(("LDG", "STG", "LDS", "STS", "LDSM", "HMMA", "ATOM", "RED", "MUFU")) else ".".join(o.split(".")[:2])
    e = int(r[iE] or 0); s = int(r[iN] or 0)
    mix[o] += e; smp[o] += s; tot += e; ts += s
nw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"total warp-instructions {tot}  per warp {tot / nw:.0f}   samples {ts}")
for o, e in mix.most_common(40):
    print(f"{o:14s} {e:10d} {100 * e / tot:5.1f} %   per-warp {e / nw:7.1f}   samples {100 * smp[o] / max(ts, 1):5.1f} %")
