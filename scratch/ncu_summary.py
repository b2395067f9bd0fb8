import csv, sys, collections
def summarize(path):
    with open(path) as f:
        lines=[l for l in f if not l.startswith('==')]
    agg=collections.OrderedDict()
    for row in csv.DictReader(lines):
        k=row['Kernel Name'].split('(')[0][-48:]
        v=float(row['Metric Value'].replace(',',''))
        agg.setdefault((k,row['Grid Size'],row['Block Size']),[]).append(v)
    tot=sum(sum(v)/len(v) for v in agg.values())
    for (k,g,b),v in agg.items():
        m=sum(v)/len(v)
        print(f"{k:50s} grid={g:14s} block={b:12s} n={len(v):3d} mean={m/1e3:9.1f} us  share={m/tot*100:5.1f}%")
summarize(sys.argv[1])
