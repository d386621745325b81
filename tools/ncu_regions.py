#!/usr/bin/env python
"""Instruction share per kernel region (line ranges of ok_kernels.cuh) from an .ncu-rep."""
import collections, csv, io, subprocess, sys
rep=sys.argv[1]
regions=[tuple(x.split(':')) for x in sys.argv[2:]]  # name:first:last
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
cs=list(csv.reader(io.StringIO(out)))
hdr=None; cur=None; per=collections.Counter(); perthr=collections.Counter()
for r in cs:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>10 and r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<10 or r[2]!='-': continue
    d=dict(zip(hdr,r))
    try: inst=int(d['Instructions Executed']); thr=int(d['Thread Instructions Executed'])
    except: continue
    per[(cur,int(r[0]))]+=inst; perthr[(cur,int(r[0]))]+=thr
tot=sum(per.values()); acc=0
for name,a,b in regions:
    a,b=int(a),int(b)
    i=sum(v for (f,l),v in per.items() if f=='ok_kernels.cuh' and a<=l<=b); t=sum(v for (f,l),v in perthr.items() if f=='ok_kernels.cuh' and a<=l<=b); acc+=i
    print(f"{name:18s} {100*i/tot:5.1f}%  lanes {t/max(i,1):5.1f}  warp-inst {i}")
i=sum(v for (f,l),v in per.items() if f=='ok_math.cuh'); t=sum(v for (f,l),v in perthr.items() if f=='ok_math.cuh'); acc+=i
print(f"ok_math.cuh        {100*i/tot:5.1f}%  lanes {t/max(i,1):5.1f}  warp-inst {i}")
print(f"other              {100*(tot-acc)/tot:5.1f}%   total {tot}")
