import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ix={n:i for i,n in enumerate(hdr)}
agg=collections.OrderedDict(); tot_all=0
for r in rows[start:]:
    if len(r)<len(hdr): continue
    v=float(r[ix['Metric Value']].replace(',',''))
    tot_all+=v
    if 'convlayer' not in r[ix['Kernel Name']]: continue
    key=(r[ix['Grid Size']],)
    a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('convlayer total ms', tot/1e6, 'all kernels ms', tot_all/1e6)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 20]:
    print(k, a[0], '%.0f'%a[1], '%.1f%%'%(100*a[1]/tot), '%.0f per launch'%(a[1]/a[0]))
