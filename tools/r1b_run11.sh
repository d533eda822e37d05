#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_kernels.py tests/test_gpu_avgpos.py -q -m gpu -x > gpurun_out/t_dist.log 2>&1; echo "dist+kernels rc=$?"
tail -5 gpurun_out/t_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b2_peer.json 2> gpurun_out/b2_peer.err; echo "bench2 rc=$?"
EVOKE_B200_OVERLAP_GATHER=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-events > gpurun_out/b2_noov.json 2> gpurun_out/b2_noov.err; echo "bench2 no-overlap rc=$?"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err
for f in b2_peer b2_noov b1; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$f", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", "eager", round(d.get("ms_per_step_eager") or 0,4), d["clocks"]["reasons"], "e2e", round(d["e2e"]["ms_per_step"],3))
    for k,v in d["kernels"].items(): print("    %-28s x%.1f  %8.1f us" % (k, v["launches_per_step"], v["avg_ms"]*1e3))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/$f.err").read()[-2500:])
PY
done
