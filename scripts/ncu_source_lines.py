"""Stall samples and executed instructions per CUDA source line of one kernel of an .ncu-rep (--import-source on).

    python scripts/ncu_source_lines.py report.ncu-rep KERNEL_ID [topn]      # KERNEL_ID: 1-based launch index
"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "launch__registers_per_thread"]
idx = [h.index(c) for c in cols if c in h]
print("launches in the report:", "  |  ".join(cols))
for r in rr[2:]:
    print("  ", "  |  ".join(r[i][:44] for i in idx))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                      f":::{kid}"], capture_output=True, text=True).stdout
cur, hdr, out = None, None, []
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != "":
        i_s, i_ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        st = {hdr[i][6:]: int(r[i] or 0) for i in range(len(hdr)) if hdr[i].startswith("stall_") and "Not Issued" not in hdr[i]}
        out.append((int(r[i_s] or 0), int(r[i_ex] or 0), cur.split("/")[-1], r[0], r[1].strip(), st))
tot, totex = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
agg = {}
for o in out:
    for k, v in o[5].items():
        agg[k] = agg.get(k, 0) + v
print(f"\nkernel #{kid}: {tot} samples, {totex} warp instructions")
print("stall totals: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 > tot))
print("\n-- by stall samples")
for s, ex, f, l, text, st in sorted(out, key=lambda o: -o[0])[:topn]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100 * s / tot:5.1f}% smp {100 * ex / totex:5.1f}% ins  {f}:{l:>4}  {text[:88]:88s} {' '.join(f'{k}={v}' for k, v in top)}")
print("\n-- by instructions executed")
for s, ex, f, l, text, st in sorted(out, key=lambda o: -o[1])[:topn]:
    print(f"{100 * ex / totex:5.1f}% ins {100 * s / tot:5.1f}% smp  {f}:{l:>4}  {text[:100]}")
