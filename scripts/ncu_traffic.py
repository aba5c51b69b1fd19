"""Per-kernel DRAM traffic / time / pipe utilisation table from an .ncu-rep captured with --set full.

    python scripts/ncu_traffic.py rep.ncu-rep out.json [trees] > table.txt

Aggregates over the captured launches of every kernel name: launches, total and mean duration, DRAM bytes read +
written per launch (dram__bytes_read.sum + dram__bytes_write.sum), DRAM / L2 / tensor-pipe utilisation.  bench.py
reads the JSON (profiles/ncu_traffic.json) for the `traffic` field of its roofline objects.
"""
import collections
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0}


def short(name):
    n = name.split("(")[0]
    for pre in ("void ", "spgnn::"):
        n = n.replace(pre, "")
    return n.split("<")[0].split("::")[-1]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    trees = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        if key not in col or r[col[key]] in ("", "n/a"):
            return None
        return float(r[col[key]].replace(",", "")) * UNIT.get(units[col[key]], 1.0)

    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = short(r[col["Kernel Name"]])
        d = agg.setdefault(k, collections.defaultdict(float))
        d["launches"] += 1
        d["time_s"] += val(r, "gpu__time_duration.sum") or 0.0
        d["dram_read"] += val(r, "dram__bytes_read.sum") or 0.0
        d["dram_write"] += val(r, "dram__bytes_write.sum") or 0.0
        for key, name in (("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
                          ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
                          ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
                          ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
                          ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct")):
            v = val(r, key)
            if v is not None:
                d[name] += v * (val(r, "gpu__time_duration.sum") or 0.0)      # time-weighted
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    # stamped with the hash of the CUDA sources the capture was taken from: bench.py ignores it once they change
    res = {"trees": trees, "csrc_sha16": bench.csrc_sha16(), "kernels": {}}
    print(f"{'kernel':34s} {'n':>3s} {'total ms':>9s} {'avg ms':>8s} {'DRAM MB/launch':>15s} {'GB/s':>7s} "
          f"{'dram%':>6s} {'l2%':>6s} {'tensor%':>8s} {'issue%':>7s}")
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["time_s"]):
        n, t = int(d["launches"]), d["time_s"]
        tr = (d["dram_read"] + d["dram_write"]) / n
        e = {"launches": n, "avg_ms": t / n * 1e3, "dram_bytes_per_launch": tr,
             "dram_read_per_launch": d["dram_read"] / n, "dram_write_per_launch": d["dram_write"] / n,
             "dram_gbs": (d["dram_read"] + d["dram_write"]) / t / 1e9 if t else None}
        for name in ("dram_pct", "l2_pct", "tensor_pct", "sm_pct", "issue_pct"):
            e[name] = d[name] / t if t and name in d else None
        res["kernels"][k] = e
        f = lambda x: f"{x:6.1f}" if x is not None else "   n/a"
        print(f"{k:34s} {n:3d} {t * 1e3:9.3f} {t / n * 1e3:8.3f} {tr / 1e6:15.1f} {e['dram_gbs'] or 0:7.0f} "
              f"{f(e['dram_pct'])} {f(e['l2_pct'])} {f(e['tensor_pct']):>8s} {f(e['issue_pct']):>7s}")
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
