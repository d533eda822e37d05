#!/bin/bash
mkdir -p gpurun_out
B="bench.py --steps 40 --warmup 5 --no-cpu-baseline"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 $B --gpus 8 > gpurun_out/b8_ov1.json 2> gpurun_out/b8_ov1.err; echo "b8 overlap rc=$?"
EVOKE_B200_OVERLAP_GATHER=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 $B --gpus 8 --no-kernel-events > gpurun_out/b8_ov0.json 2> gpurun_out/b8_ov0.err; echo "b8 no-overlap rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 $B --gpus 4 --no-kernel-events > gpurun_out/b4_ov1.json 2> gpurun_out/b4_ov1.err; echo "b4 overlap rc=$?"
EVOKE_B200_OVERLAP_GATHER=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 $B --gpus 4 --no-kernel-events > gpurun_out/b4_ov0.json 2> gpurun_out/b4_ov0.err; echo "b4 no-overlap rc=$?"
for f in b8_ov1 b8_ov0 b4_ov1 b4_ov0; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$f.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$f", round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,3), "Mpairs/s", "eager", round(d.get("ms_per_step_eager") or 0,4), d["clocks"]["reasons"], "e2e", round(d["e2e"]["ms_per_step"],3), "loss", d["loss"])
    for k,v in d["kernels"].items(): print("    %-28s x%.1f  %8.1f us" % (k, v["launches_per_step"], v["avg_ms"]*1e3))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/$f.err").read()[-2500:])
PY
done
