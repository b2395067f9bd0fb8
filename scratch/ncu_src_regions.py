"""Aggregate an `ncu --page source --csv` dump by address region and stall reason.
usage: ncu_src_regions.py file.csv [boundaries as hex offsets from the kernel start, ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
ia, isrc, iall = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
iex = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ia], 16)
bounds = [int(x, 16) for x in sys.argv[2:]] or [0]
bounds = sorted(set([0] + bounds))
agg = {b: {"n": 0, "samples": 0, "exec": 0, "st": {}} for b in bounds}
for r in data:
    off = int(r[ia], 16) - base
    b = max(x for x in bounds if x <= off)
    a = agg[b]
    a["n"] += 1
    a["samples"] += int(r[iall] or 0)
    a["exec"] += 1 if int(r[iex] or 0) > 0 else 0
    for i in stalls:
        v = int(r[i] or 0)
        if v:
            a["st"][hdr[i]] = a["st"].get(hdr[i], 0) + v
tot = sum(a["samples"] for a in agg.values())
print(f"total samples {tot}")
for b in bounds:
    a = agg[b]
    top = sorted(a["st"].items(), key=lambda kv: -kv[1])[:6]
    print(f"region 0x{b:05x}: {a['n']:5d} instrs ({a['exec']} executed) samples {a['samples']:7d} ({100 * a['samples'] / max(tot, 1):5.1f}%)  " +
          ", ".join(f"{k[6:]} {v}" for k, v in top))
