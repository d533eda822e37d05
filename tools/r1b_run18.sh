#!/bin/bash
mkdir -p gpurun_out
B="bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-kernel-events"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 $B --gpus 8 > gpurun_out/final_bench_n8.json 2> gpurun_out/final_bench_n8.err; echo "b8 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 $B --gpus 4 > gpurun_out/final_bench_n4.json 2> gpurun_out/final_bench_n4.err; echo "b4 rc=$?"
for f in final_bench_n8 final_bench_n4; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$f", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", d["clocks"], "e2e", round(d["e2e"]["ms_per_step"],3), "loss", d["loss"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/$f.err").read()[-2000:])
PY
done
