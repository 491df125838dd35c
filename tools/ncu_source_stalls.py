import csv,sys
from collections import defaultdict
rows=list(csv.reader(open(sys.argv[1])))
which=int(sys.argv[2]) if len(sys.argv)>2 else 0
blocks=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name":
        cur={'name':r[1],'hdr':None,'data':[]}; blocks.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r; continue
    if len(r)==len(cur['hdr']): cur['data'].append(r)
print([ (b['name'][:60], len(b['data'])) for b in blocks])
b=blocks[which]; hdr=b['hdr']; data=b['data']; ix={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[ix['# Samples']]) for r in data)
print("total samples",tot)
agg=defaultdict(lambda:[0,0,0])
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
st_tot=defaultdict(int)
for r in data:
    toks=r[ix['Source']].split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    op=op.split('.')[0]
    agg[op][0]+=int(r[ix['# Samples']]); agg[op][1]+=int(r[ix['Instructions Executed']]); agg[op][2]+=1
    for h in stalls: st_tot[h]+=int(r[ix[h]])
for op,(s,e,n) in sorted(agg.items(), key=lambda x:-x[1][0])[:14]:
    print(f"{op:12s} samples {s:7d} ({100*s/tot:5.1f}%) executed {e:10d} static {n}")
print({k:v for k,v in sorted(st_tot.items(), key=lambda x:-x[1])[:8]})
print("hottest:")
for r in sorted(data, key=lambda r:-int(r[ix['# Samples']]))[:int(sys.argv[3]) if len(sys.argv)>3 else 20]:
    print(r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']][:80], {h[6:]:r[ix[h]] for h in stalls if int(r[ix[h]])>20})
print("per-op stall breakdown")
per=defaultdict(lambda: defaultdict(int))
for r in data:
    toks=r[ix['Source']].split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    op=op.split('.')[0]
    for h in stalls: per[op][h[6:]]+=int(r[ix[h]])
for op in ['FFMA2','BRA','ISETP','IADD3','MOV','IMAD','PRMT','SHF','LEA','LDS','LDG','STG']:
    print(op, dict(sorted(per[op].items(), key=lambda x:-x[1])[:5]))
