#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for mode in peer sym; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --shard-mode $mode > gpurun_out/b8_$mode.json 2> gpurun_out/b8_$mode.err; echo "bench8 $mode rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline --shard-mode peer > gpurun_out/b4_peer.json 2> gpurun_out/b4_peer.err; echo "bench4 rc=$?"
for f in b8_peer b8_sym b4_peer; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$f", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", "eager", round(d.get("ms_per_step_eager") or 0,4), d["clocks"]["reasons"], "e2e", round(d["e2e"]["ms_per_step"],3))
    for k,v in d["kernels"].items(): print("    %-28s x%.1f  %8.1f us" % (k, v["launches_per_step"], v["avg_ms"]*1e3))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/$f.err").read()[-2500:])
PY
done
