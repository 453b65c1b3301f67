import csv, subprocess, collections, sys
rep=sys.argv[1]; nw=8192.0
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hidx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hidx[0]]; body=rows[hidx[0]+1:(hidx[1]-1 if len(hidx)>1 else len(rows))]
col={n:i for i,n in enumerate(h)}
print(h)
def f(r,n):
    try: return float(r[col[n]])
    except: return 0.0
buckets=collections.defaultdict(lambda: [0,0,collections.Counter()])
for r in body:
    e=f(r,'Instructions Executed')/nw
    b = 'prologue(<30/CTA)' if e<30 else ('task(30-140)' if e<140 else 'inner(>140)')
    buckets[b][0]+=e; buckets[b][1]+=f(r,'# Samples')
    srcs=r[col['Source']].split()
    if srcs:
        o=srcs[1] if srcs[0].startswith('@') and len(srcs)>1 else srcs[0]
        buckets[b][2][o.split('.')[0]]+=e
for b,(e,s,ops) in buckets.items():
    print(b, "instr/CTA %.0f samples %d"%(e,s)); print("   ", [(k,int(v)) for k,v in ops.most_common(25)])
