#!/bin/bash
# dev loop on the GPU box: parity + ncu kernel durations of the fused CP-conv rows kernels.  usage: tools/cprows_run.sh <outdir> [cases...]
out=gpurun_out/$1; shift
mkdir -p $out
cases=${@:-fwd1 bwd1 fwd2}
for c in $cases; do
  timeout 120 python tools/cprows_check.py --case $c > $out/$c.log 2>&1; echo "$c rc=$?"; grep err_ $out/$c.log
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cpconv_rows --csv --log-file $out/launches_$c.csv python tools/cprows_check.py --case $c --steps 2 > $out/ncu_$c.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('$out/launches_$c.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
import collections
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:45]].append(float(r[vi])/1e3)
for k,v in d.items():
    big=[x for x in v if x>max(v)*0.5]
    print(k, 'n=%d'%len(big), 'avg us %.1f'%(sum(big)/len(big)), 'min %.1f'%min(big))
PY
done
