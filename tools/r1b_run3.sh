#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_estrip.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/t_all.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
EVK_STREAMK=0 $B > gpurun_out/b_sk0.json 2> gpurun_out/b_sk0.err
EVK_STREAMK=1 $B > gpurun_out/b_sk1.json 2> gpurun_out/b_sk1.err
for f in sk0 sk1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", "eager", round(d.get("ms_per_step_eager") or 0,4), {k:round(v["avg_ms"],4) for k,v in d["kernels"].items()}, d["clocks"], "e2e", round(d["e2e"]["ms_per_step"],3))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/b_$f.err").read()[-1500:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-clocks > gpurun_out/ncu_r1b.log 2>&1; echo "ncu rc=$?"
