"""Executed warp-instruction totals per SASS opcode for one kernel of an .ncu-rep (source page)."""
import csv, subprocess, sys, collections
rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
blocks, cur = [], None
for r in rows:
    if "Source" in r and "# Samples" in r:
        cur = {"h": r, "rows": []}; blocks.append(cur)
    elif cur is not None and len(r) == len(cur["h"]):
        cur["rows"].append(r)
b = blocks[which]; h = b["h"]
i_src, i_ex, i_s = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, smp = collections.Counter(), collections.Counter()
tot = 0
for r in b["rows"]:
    try:
        ex = int(r[i_ex] or 0); s_ = int(r[i_s] or 0)
    except ValueError:
        continue
    t = r[i_src].strip().split()
    if not t: continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    op = op.split(".")[0]
    ops[op] += ex; smp[op] += s_; tot += ex
print(f"kernel #{which}: {tot/1e6:.1f} M warp instructions")
for op, n in ops.most_common(22):
    print(f"  {op:10s} {n/1e6:9.2f} M  {100*n/tot:5.1f}%   samples {100*smp[op]/max(1,sum(smp.values())):5.1f}%")
