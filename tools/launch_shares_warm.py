import csv, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = []
cur = {}
for row in csv.DictReader(lines):
    key=(row['ID'])
    cur.setdefault(key, {'name': row['Kernel Name'].split('(')[0], 'grid': row['Grid Size']})
    v = float(row['Metric Value'].replace(',', '')); unit=row['Metric Unit']; m=row['Metric Name']
    if m=='gpu__time_duration.sum':
        if unit in ('usecond','us'): v*=1e3
        elif unit in ('msecond','ms'): v*=1e6
        cur[key]['t']=v
    else:
        mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(unit,1)
        cur[key][m]=v*mult
rows=[cur[k] for k in sorted(cur, key=lambda x:int(x))]
names=[r['name'] for r in rows]
starts=[i for i,n in enumerate(names) if 'k_resize_roi_swap' in n]
ends=[i for i,n in enumerate(names) if 'k_post' in n]
s0=starts[1]; e0=[e for e in ends if e>s0][0]
call=rows[s0:e0+1]
tot=sum(r['t'] for r in call)
agg=collections.defaultdict(lambda:[0,0.0,0.0,0.0])
for r in call:
    a=agg[r['name']]; a[0]+=1; a[1]+=r['t']; a[2]+=r.get('dram__bytes_read.sum',0)+r.get('dram__bytes_write.sum',0); a[3]+=r.get('lts__t_bytes.sum',0)
print(f'{len(call)} launches in one call; total {tot/1e3:.1f} us')
for n,(c,t,d,l) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f'{n:42s} x{c:3d} {t/1e3:9.1f} us {100*t/tot:5.1f}%  dram {d/1e6:8.1f} MB  L2 {l/1e6:8.1f} MB  L2 GB/s {l/t:7.0f}')
if len(sys.argv)>2:
    for i,r in enumerate(call): print(i, r['name'][:34], f"{r['t']/1e3:.1f}", r['grid'])
