#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_train.py > gpurun_out/train_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches.csv python scripts/prof_train.py > gpurun_out/ncu_train.log 2>&1
echo "ncu train exit $?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/train_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki,vi=h.index('Kernel Name'),h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hi+2:]:
    if len(r)<=vi: continue
    a=agg.setdefault(r[ki][:80],[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:30]:
    print('%-82s %4d %9.1f us %5.1f%%'%(k,a[0],a[1]/1e3,100*a[1]/tot))
print('total us', tot/1e3)
PY
