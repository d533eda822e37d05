#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-clocks"
for v in 0 1 2; do
  export EVK_K3_VARIANT=$v
  timeout 300 python -m pytest tests/test_gpu_estrip.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/t_v$v.log 2>&1; echo "variant $v tests rc=$?"; tail -1 gpurun_out/t_v$v.log
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_kernel' -c 24 --csv --log-file gpurun_out/launches_v$v.csv $CMD > gpurun_out/ncu_v$v.log 2>&1
  python - <<PY
import csv,collections
lines=[l for l in open('gpurun_out/launches_v$v.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row['Kernel Name'][:70],[]).append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items(): print("  variant $v", f"{k:60s} n={len(v):3d} avg={sum(v)/len(v)/1000:9.1f} us  min={min(v)/1000:8.1f}")
PY
done
for v in 0 2; do
  EVK_K3_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-kernel-events > gpurun_out/b1_v$v.json 2> gpurun_out/b1_v$v.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/b1_v$v.json').read().strip().splitlines() if l.startswith('{')][-1])
print('variant $v graph step', round(d['ms_per_step'],4), 'ms', d['clocks'])"
done
