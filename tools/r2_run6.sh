# 8 GPUs: world-8 parity tests, then the sharded bench with A/B variants, then cfg4
python -m pytest tests/test_gpu_distributed.py -m gpu -q -k "(sharded_equals_oracle and 8 and (peer or rs)) or (sharded_mpc and mixed and 8 and bf16)" > gpurun_out/r2_t6_dist8.log 2>&1; echo "dist rc=$?" >> gpurun_out/r2_t6_dist8.log; tail -6 gpurun_out/r2_t6_dist8.log
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager ${@:3} > gpurun_out/$2 2>> gpurun_out/r2_t6_bench.err; echo "$2 rc=$?"; }
run 29621 r2_t6_n8_default.json
EVOKE_B200_FOLDED_SYNC=0 run 29622 r2_t6_n8_barriers.json --no-kernel-events --no-dropin --no-sustained
EVOKE_B200_SCATTER_ROTATE=0 run 29623 r2_t6_n8_norotate.json --no-kernel-events --no-dropin --no-sustained
EVOKE_B200_MASK_FREE=0 run 29624 r2_t6_n8_mask.json --no-kernel-events --no-dropin --no-sustained
run 29625 r2_t6_n8_cfg4.json --config cfg4 --no-dropin
nvidia-smi nvlink -gt d 2>/dev/null | head -40 > gpurun_out/r2_t6_nvlink.txt
tail -c 600 gpurun_out/r2_t6_bench.err
