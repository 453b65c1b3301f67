#!/usr/bin/env python
"""Instruction share per contiguous SASS segment (same execution count) of a profiled kernel."""
import csv, subprocess, collections, sys
rep=sys.argv[1]; nw=float(sys.argv[2]) if len(sys.argv)>2 else 8192.0
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hidx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hidx[0]]; body=rows[hidx[0]+1:(hidx[1]-1 if len(hidx)>1 else len(rows))]
col={n:i for i,n in enumerate(h)}
def f(r,n):
    try: return float(r[col[n]])
    except: return 0.0
toti=sum(f(r,'Instructions Executed') for r in body); tots=sum(f(r,'# Samples') for r in body)
segs=[];cur=None
for i,r in enumerate(body):
    e=f(r,'Instructions Executed')
    if cur is None or abs(e-cur['lvl'])>0.05*max(e,cur['lvl'],1):
        cur={'lvl':e,'s':i,'n':0,'inst':0,'smp':0,'ops':collections.Counter()}; segs.append(cur)
    cur['n']+=1; cur['inst']+=e; cur['e']=i; cur['smp']+=f(r,'# Samples')
    srcs=r[col['Source']].split()
    if srcs:
        o=srcs[1] if srcs[0].startswith('@') and len(srcs)>1 else srcs[0]
        cur['ops'][o.split('.')[0]]+=1
print("total warp instr %.4g, per CTA %.0f"%(toti, toti/nw))
for sg in sorted(segs,key=lambda s:-s['inst'])[:int(sys.argv[3]) if len(sys.argv)>3 else 24]:
    print("rows %5d-%5d n=%4d execs/CTA=%7.1f inst=%5.1f%% smp=%5.1f%%  %s"%(sg['s'],sg['e'],sg['n'],sg['lvl']/nw,100*sg['inst']/toti,100*sg['smp']/tots, dict(sg['ops'].most_common(6))))
