import sys
rows=[tuple(map(int,l.split())) for l in open(sys.argv[1])]
t0=min(r[0] for r in rows); t1=max(r[1] for r in rows)
print("blocks", len(rows), "span ms", (t1-t0)/1e6, "mean block ms", sum(r[1]-r[0] for r in rows)/len(rows)/1e6)
# average concurrency = sum of durations / span
print("avg resident blocks", sum(r[1]-r[0] for r in rows)/(t1-t0), "per SM", sum(r[1]-r[0] for r in rows)/(t1-t0)/148)
# concurrency histogram over time (1000 bins) in the middle 60%
import bisect
ev=sorted([(r[0],1) for r in rows]+[(r[1],-1) for r in rows])
cur=0; last=t0; acc={}
for t,d in ev:
    if t0+0.2*(t1-t0) < t < t0+0.8*(t1-t0):
        acc[cur]=acc.get(cur,0)+(t-last)
    last=t; cur+=d
tot=sum(acc.values())
ks=sorted(acc)
import itertools
c=0
for q in (0.05,0.25,0.5,0.75,0.95):
    s=0
    for k in ks:
        s+=acc[k]
        if s>=q*tot:
            print("resident-blocks quantile",q,k); break
sm={}
for r in rows: sm[r[2]]=sm.get(r[2],0)+1
print("SMs used", len(sm), "min/max blocks per SM", min(sm.values()), max(sm.values()))
