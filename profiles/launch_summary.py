import csv, re, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
agg=collections.OrderedDict(); tot=0
for row in csv.DictReader(lines):
    v=float(row['Metric Value']); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    k=re.sub(r'.*::','',re.sub(r'\(.*','',row['Kernel Name']))
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
print('total us',round(tot))
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:45s} n={n:3d} {t:9.1f} us avg {t/n:8.1f} {100*t/tot:5.1f}%")
