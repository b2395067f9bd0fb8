"""Per-source-line totals (instructions executed, stall samples) of one kernel in an `ncu --set full --import-source on` report.
    python scratch/ncu_lines.py report.ncu-rep kernel_regex [top] [launch index]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}', '--print-source', 'sass,cuda',
                      '--launch-count', '1'] , capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: [0, 0, ''])
fname, hdr = None, None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); continue
    if hdr is None: continue
    try:
        line = int(r[0]); ex = int(r[ie] or 0); sm = int(r[isamp] or 0)
    except Exception:
        continue
    a = agg[(fname, line)]
    a[0] += ex; a[1] += sm; a[2] = r[1]
tot_e = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f'total warp instructions {tot_e}, samples {tot_s}')
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{a[0] / tot_e * 100:5.1f}% instr {a[1] / max(tot_s, 1) * 100:5.1f}% samp  {f}:{l:<5d} {a[2].strip()[:110]}')
