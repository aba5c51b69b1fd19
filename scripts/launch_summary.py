"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals of one training step.

    python scripts/launch_summary.py gpurun_out/launches.csv [marker_kernel_substring]
"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
iK, iV, iU, iG, iB = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size"), hdr.index("Block Size")
L = []
for r in rows[1:]:
    try:
        v = float(r[iV].replace(",", ""))
    except ValueError:
        continue
    u = r[iU]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    L.append((r[iK].split("(")[0][-70:], ms, r[iG], r[iB]))
marker = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "sgd_"
ends = [i for i, x in enumerate(L) if marker in x[0]]
if len(ends) >= 2:
    # a training step = the launches between two optimiser updates; bench.py also runs inference passes between its
    # timed steps and its profile step, so take the segment with the fewest launches (a pure training step)
    segs = [L[a + 1: b + 1] for a, b in zip(ends[:-1], ends[1:])]
    step = min(segs, key=len)
    print(f"one training step: launches {len(step)}, sum {sum(x[1] for x in step):.2f} ms "
          f"({len(ends)} optimiser updates in the capture)")
else:
    step = L
agg = collections.OrderedDict()
for k, ms, g, b in step:
    d = agg.setdefault(k, [0, 0.0])
    d[0] += 1; d[1] += ms
tot = sum(d[1] for d in agg.values())
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:9.3f} ms {100*ms/tot:5.1f}%  x{n:3d}  {k}")
if "--list" in sys.argv:
    for k, ms, g, b in step:
        print(f"   {ms:8.3f}  {g:>14s} {b:>12s}  {k}")
