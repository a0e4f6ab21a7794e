#!/usr/bin/env python
"""tools/ncu_regions.py REP: where the issue slots of a kernel go, straight from the source page of an ncu report
(no line map to maintain): consecutive SASS instructions with the same execution count are one region; prints its
share of all warp instructions, the lanes active in it and its share of the stall samples."""
import csv, io, subprocess, sys
rep=sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hi[0]]
ia, ie, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ist = hdr.index("Warp Stall Sampling (All Samples)")
d = [r for r in rows[hi[0] + 1:] if len(r) == len(hdr) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in d); tots=sum(int(r[ist] or 0) for r in d)
print("total", tot, "static", len(d))
# group consecutive instructions with same exec count (within 2%)
groups=[]
for k,r in enumerate(d):
    e,t,s=int(r[ie]),int(r[it]),int(r[ist] or 0)
    if groups and abs(groups[-1]['e']-e)<=0.02*max(e,1) and e>0:
        g=groups[-1]; g['n']+=1; g['sum']+=e; g['thr']+=t; g['st']+=s; g['last']=k
    else:
        groups.append(dict(first=k,last=k,e=e,n=1,sum=e,thr=t,st=s,src=r[ia].strip()))
for g in groups:
    if g['sum']>tot*0.002:
        print(f"{g['first']:5d}-{g['last']:5d} n={g['n']:4d} exec/inst={g['e']/1e3:8.1f}K share={100*g['sum']/tot:5.2f}% lanes={g['thr']/max(g['sum'],1):5.1f} stall={100*g['st']/max(tots,1):5.2f}%  {g['src'][:60]}")
