import csv,subprocess,collections,sys
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 50
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
for i,r in enumerate(rows):
    if r and r[0]=='Line No': h=i; break
hdr=rows[h]; ia=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
per={}; tot=0; tsamp=0
for r in rows[h+1:]:
    if len(r)!=len(hdr) or not r[0].strip().isdigit(): continue
    try: c=int(r[ia]); s=int(r[isamp])
    except: continue
    k=(int(r[0]), r[1].strip()[:120])
    v=per.setdefault(k,[0,0]); v[0]+=c; v[1]+=s; tot+=c; tsamp+=s
print('total inst',tot,'samples',tsamp)
for k,v in sorted(per.items(), key=lambda kv:-kv[1][0])[:topn]:
    print(f'{k[0]:5d} {v[0]:9d} {100*v[0]/tot:5.1f}% samp {100*v[1]/max(tsamp,1):5.1f}% | {k[1]}')
