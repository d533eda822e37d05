#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_estrip.py -q -m gpu > gpurun_out/t_estrip.log 2>&1; echo "estrip rc=$?"
tail -8 gpurun_out/t_estrip.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -m gpu > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -4 gpurun_out/t_parity.log
python tools/overlap_probe.py > gpurun_out/overlap_probe.log 2>&1; cat gpurun_out/overlap_probe.log | tail -12
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b_strip1.json 2> gpurun_out/b_strip1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b_strip1.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", {k:round(v["avg_ms"],4) for k,v in d["kernels"].items()}, d["clocks"], d["e2e"]["ms_per_step"])
PY
