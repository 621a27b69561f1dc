import sys,csv
for path in sys.argv[1:]:
    rows=[r for r in csv.reader(open(path,errors='replace')) if r]
    hi=next(i for i,r in enumerate(rows) if 'Metric Name' in r)
    h=rows[hi]; m=h.index('Metric Name'); v=h.index('Metric Value'); u=h.index('Metric Unit')
    tot={}
    for r in rows[hi+1:]:
        if len(r)>v:
            sc={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'ns':1e-9,'us':1e-6,'ms':1e-3,'nsecond':1e-9,'usecond':1e-6,'msecond':1e-3,'second':1}.get(r[u],1)
            tot[r[m]]=tot.get(r[m],0)+float(r[v].replace(',',''))*sc
    print(path, {k:'%.3g'%x for k,x in tot.items()})
