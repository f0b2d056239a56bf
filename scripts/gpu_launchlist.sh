#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0"
# warm-up = 3 steps (~120 launches each incl. torch RNG/copies); capture one full step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 130 --csv \
    --log-file gpurun_out/launches2.csv $CMD > gpurun_out/ncu_launches2.log 2>&1
echo "launch list exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches2.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict(); seq=[]
for r in rows:
    name=r[4].split('(')[0].split('::')[-1][:40]; us=float(r[-1])/1000.0 if r[-2]=='ns' else float(r[-1])
    seq.append((name,us)); agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=us
tot=sum(v[1] for v in agg.values())
print('total us', round(tot,1), 'launches', len(seq))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f'{k:42s} n={v[0]:3d} total {v[1]:8.1f} us  avg {v[1]/v[0]:7.1f}  {100*v[1]/tot:5.1f}%')
print([ (n[:12], round(u,1)) for n,u in seq[:60]])
PY
