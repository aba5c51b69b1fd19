"""Static SASS opcode census of the built library: per kernel, how many tcgen05 / TMA / mbarrier / memory instructions
the sm_100a code holds (cuobjdump -sass; no GPU needed).

    python scripts/sass_opcodes.py [spgnn_b200/libspgnn_b200.so] > profiles/rNN_sass_opcodes.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spgnn_b200", "libspgnn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL",
       "MUFU", "ATOMS", "ATOMG", "RED", "BAR", "FFMA", "HMMA", "LDL", "STL"]
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = kernels.setdefault(re.sub(r"\(.*", "", name)[-64:], collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["_total"] += 1
        op = m.group(1)
        for k in KEY:
            if op.startswith(k):
                cur[k] += 1
                break
print(f"{'kernel':66s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEY))
for name, c in sorted(kernels.items(), key=lambda kv: -kv[1]["_total"]):
    print(f"{name:66s} {c['_total']:6d} " + " ".join(f"{c[k]:7d}" if c[k] else f"{'.':>7s}" for k in KEY))
tot = collections.Counter()
for c in kernels.values():
    tot.update(c)
print(f"{'ALL KERNELS (' + str(len(kernels)) + ')':66s} {tot['_total']:6d} " + " ".join(f"{tot[k]:7d}" for k in KEY))
