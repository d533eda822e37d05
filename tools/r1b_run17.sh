#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -q -m gpu > gpurun_out/t_dist.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/t_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "bench2 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/final_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
print(round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", d["clocks"], "e2e", round(d["e2e"]["ms_per_step"],3), d["config"]["parallelism"])
for k,v in d["kernels"].items(): print("    %-28s x%.1f  %8.1f us" % (k, v["launches_per_step"], v["avg_ms"]*1e3))
PY
