import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
d=collections.OrderedDict()
for r in rows[hdr+2:]:
    if len(r)>vi: d.setdefault(r[ki][:36],[]).append(float(r[vi].replace(',','')))
for k,v in d.items(): print(f"{k:38s} n={len(v):3d} median={sorted(v)[len(v)//2]/1e3:8.1f} us  min={min(v)/1e3:8.1f}")
