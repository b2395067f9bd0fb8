import csv, sys, subprocess, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
mode = sys.argv[4] if len(sys.argv) > 4 else 'sass'
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name',f'regex:{kern}','--print-source',mode],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hi=[i for i,r in enumerate(rows) if r and r[0] in ('Address','#','Line')]
hi=hi[0]
hdr=rows[hi]; idx={h:i for i,h in enumerate(hdr)}
col=idx['Warp Stall Sampling (All Samples)']; src=idx['Source']; ie=idx.get('Instructions Executed')
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data=[]
for n,r in enumerate(rows[hi+1:]):
    try: data.append((int(r[col]), n, r))
    except: pass
tot=sum(d[0] for d in data); print('total samples',tot, 'lines', len(data))
agg={s:0 for s in stall_cols}
for c,n,r in data:
    for s in stall_cols:
        try: agg[s]+=int(r[idx[s]])
        except: pass
print('stall mix:', ', '.join(f"{k[6:]}={v/tot*100:.1f}%" for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]))
for c,n,r in sorted(data,reverse=True)[:top]:
    st=max(stall_cols,key=lambda s:int(r[idx[s]] or 0))
    print(f"{c:6d} {c/tot*100:5.1f}%  #{n:5d} ex={r[ie] if ie else '':>8s} {st[6:]:10s} {r[src][:110]}")
