# 8 GPUs: side-by-side contractions A/B
python -m pytest tests/test_gpu_distributed.py -m gpu -q -k "sharded_equals_oracle and peer and 8 and bf16 and not fp32" > gpurun_out/r2_t8_dist8.log 2>&1; echo "dist rc=$?" >> gpurun_out/r2_t8_dist8.log; tail -3 gpurun_out/r2_t8_dist8.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager --no-kernel-events ${@:3} > gpurun_out/$2 2>> gpurun_out/r2_t8_bench.err; echo "$2 rc=$?"; }
run 29631 r2_t8_n8_sbs74.json
EVOKE_B200_SIDE_BY_SIDE_CTAS=0 run 29632 r2_t8_n8_sbs0.json --no-dropin --no-sustained
EVOKE_B200_SIDE_BY_SIDE_CTAS=60 run 29633 r2_t8_n8_sbs60.json --no-dropin --no-sustained
tail -c 400 gpurun_out/r2_t8_bench.err
