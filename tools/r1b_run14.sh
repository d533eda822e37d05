#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/t_all.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b1.json").read().strip().splitlines() if l.startswith("{")][-1])
print(round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", "eager", round(d.get("ms_per_step_eager") or 0,4), d["clocks"], "e2e", round(d["e2e"]["ms_per_step"],3))
for k,v in d["kernels"].items(): print("    %-28s x%.1f  %8.1f us" % (k, v["launches_per_step"], v["avg_ms"]*1e3))
print(d["roofline"]); print(d["roofline_step"])
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-clocks"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_kernel|l2norm_bwd|stats_fused|posmask' -c 60 --csv --log-file gpurun_out/launches14.csv $CMD > gpurun_out/ncu14.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv,collections
lines=[l for l in open('gpurun_out/launches14.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row['Kernel Name'][:70],[]).append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items(): print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v)/1000:9.1f} us")
PY
