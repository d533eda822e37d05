#!/bin/bash
# bring-up of the E-strip backward: tests, then A/B benches
mkdir -p gpurun_out
python -m pytest tests/test_gpu_estrip.py -x -q -m gpu > gpurun_out/t_estrip.log 2>&1; echo "estrip rc=$?"
tail -5 gpurun_out/t_estrip.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/t_parity.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
EVOKE_B200_ESTRIP=0 $B > gpurun_out/b_strip0.json 2> gpurun_out/b_strip0.err
EVOKE_B200_ESTRIP=1 EVOKE_B200_BWD_CHUNKS=1 $B > gpurun_out/b_strip1_c1.json 2> gpurun_out/b_strip1_c1.err
EVOKE_B200_ESTRIP=1 EVOKE_B200_BWD_CHUNKS=4 $B > gpurun_out/b_strip1_c4.json 2> gpurun_out/b_strip1_c4.err
EVOKE_B200_ESTRIP=1 EVOKE_B200_BWD_CHUNKS=8 $B > gpurun_out/b_strip1_c8.json 2> gpurun_out/b_strip1_c8.err
for f in strip0 strip1_c1 strip1_c4 strip1_c8; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", "eager", d.get("ms_per_step_eager"), {k:round(v["avg_ms"],4) for k,v in d["kernels"].items()}, d["clocks"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/b_$f.err").read()[-1500:])
PY
done
