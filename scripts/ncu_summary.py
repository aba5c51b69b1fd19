"""Summarise an .ncu-rep: key raw metrics + top stalled SASS lines (needs ncu on PATH; run in the authoring container)."""
import csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:100], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warp_latency_per_inst_issued.ratio", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    u = dict(zip(hdr, units))
    for k in keys:
        if k in d:
            print(f"  {k:80s} {d[k]:>16s} {u[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# one block per kernel: a title row, a header row ("Source", "# Samples", ...), then one row per SASS instruction
blocks, cur = [], None
for r in rows:
    if "Source" in r and "# Samples" in r:
        cur = {"h": r, "rows": []}
        blocks.append(cur)
    elif cur is not None and len(r) == len(cur["h"]):
        cur["rows"].append(r)
for bi, b in enumerate(blocks):
    h = b["h"]
    i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    data = []
    for n, r in enumerate(b["rows"]):
        try:
            data.append((int(r[i_s] or 0), r[i_src].strip(), int(r[i_ex] or 0), n))
        except ValueError:
            pass
    tot = sum(x[0] for x in data) or 1
    print(f"  -- kernel #{bi}: {len(data)} SASS instructions, {tot} samples; top by stall samples:")
    for s_, src_, ex, n in sorted(data, key=lambda x: -x[0])[:topn]:
        print(f"  {100*s_/tot:5.1f}%  ex={ex:10d} #{n:5d} {src_[:100]}")
