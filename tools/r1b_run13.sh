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
