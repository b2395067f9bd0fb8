"""profiles/sass_summary.txt: opcode histogram of the shipped library (cuobjdump -sass), per kernel, for the mnemonics that
prove which hardware path a kernel takes (tcgen05 = UTCHMMA/UTCBAR/LDTM, TMA = UTMALDG/UTMASTG, native shared atomics =
ATOMS.ADD vs the CAS loop ATOMS.CAST.SPIN, cp.async = LDGSTS, cluster = UCGABAR/.2CTA).
    python scratch/sass_summary.py > profiles/sass_summary.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "multimodal_survival_prediction_b200/libb200surv.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "ATOMS.ADD", "ATOMS.CAST.SPIN", "ATOMS.POPC.INC",
        "ATOMG", "RED.E", "LDGSTS", "LDG.E.128", "STG.E.128", "LDS.128", "MATCH.ANY", "SHFL", "VOTE", "DADD", "DFMA", "MUFU.EX2", "UCGABAR", "BAR.SYNC", "MEMBAR"]
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for p in pats:
            if op == p or op.startswith(p + ".") or (p == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                per[cur][p] += 1
dem = subprocess.run(["cu++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"SASS opcode summary of {so} (cuobjdump -sass, sm_100a); {len(per)} kernels")
print("per kernel: total instructions, then the non-zero counts of the marker mnemonics\n")
for (k, c), d in zip(per.items(), dem):
    name = d.replace("b200surv::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"^void ", "", name)
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):          # cut the parameter list: first "(" outside template brackets
        depth += ch == "<"; depth -= ch == ">"
        if ch == "(" and depth == 0:
            cut = i; break
    name = name[:cut][:92]
    marks = " ".join(f"{p}={c[p]}" for p in pats if c[p])
    print(f"{name:92s} {c['_total']:6d}  {marks}")
    tot.update(c)
print("\nlibrary totals: " + " ".join(f"{p}={tot[p]}" for p in pats if tot[p]) + f" instructions={tot['_total']}")
