"""Per-kernel stall-reason totals and top SASS lines (with their dominant stall reason) from an .ncu-rep source page.

    python scripts/ncu_stalls.py report.ncu-rep [kernel_index] [topn]
"""
import csv, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
blocks, cur = [], None
for r in rows:
    if "Source" in r and "# Samples" in r:
        cur = {"h": r, "rows": []}
        blocks.append(cur)
    elif cur is not None and len(r) == len(cur["h"]):
        cur["rows"].append(r)
for bi, b in enumerate(blocks):
    if which is not None and bi != which:
        continue
    h = b["h"]
    i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n: 0 for _, n in stall_cols}
    data = []
    for n, r in enumerate(b["rows"]):
        try:
            smp = int(r[i_s] or 0)
        except ValueError:
            continue
        st = {nm: int(r[i] or 0) for i, nm in stall_cols}
        for k, v in st.items():
            tot[k] += v
        data.append((smp, r[i_src].strip(), int(r[i_ex] or 0), n, st))
    total = sum(x[0] for x in data) or 1
    print(f"== kernel #{bi}: {len(data)} SASS, {total} samples")
    print("   stall totals: " + ", ".join(f"{k[6:]} {100*v/total:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 200 > total))
    for smp, s_, ex, n, st in sorted(data, key=lambda x: -x[0])[:topn]:
        top = max(st.items(), key=lambda kv: kv[1])
        print(f"  {100*smp/total:5.1f}%  ex={ex:10d} #{n:5d} {top[0][6:]:>12s}  {s_[:90]}")
