#!/bin/bash
# per-kernel breakdown of one training step and of MU iterations (ncu launch lists)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/t64_train.csv python scripts/train_time.py > gpurun_out/t64_train.log 2>&1
echo "train rc=$?"
MU_FRAMES=225000 MU_ITERS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t64_mu.csv python scripts/mu_scaling.py > gpurun_out/t64_mu.log 2>&1
echo "mu rc=$?"
python - <<'PY'
import csv,collections
for name in ("train","mu"):
    rows=[r for r in csv.DictReader(l for l in open('gpurun_out/t64_%s.csv'%name) if not l.startswith('=='))]
    agg=collections.OrderedDict()
    for r in rows:
        n=r['Kernel Name'].split('(')[0][-52:]
        v=float(r['Metric Value'].replace(',',''))
        u=r['Metric Unit']
        ms=v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
        a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=ms
    tot=sum(a[1] for a in agg.values())
    print("==",name,"total %.2f ms over %d launches"%(tot,len(rows)))
    for n,(c,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
        print("%-54s %5d x %8.3f ms = %9.3f ms %5.1f%%"%(n,c,ms/c,ms,100*ms/tot))
PY
