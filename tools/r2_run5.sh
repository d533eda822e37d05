# 2 GPUs: the late-rank test again, then A/B benches of the sharded step
python -m pytest tests/test_gpu_distributed.py -m gpu -q -x -k "misses_a_barrier or sharded_mpc" > gpurun_out/r2_t5_dist.log 2>&1; echo "dist rc=$?" >> gpurun_out/r2_t5_dist.log; tail -5 gpurun_out/r2_t5_dist.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > gpurun_out/$2 2>> gpurun_out/r2_t5_bench.err; echo "$2 rc=$?"; }
run 29611 r2_t5_n2_default.json
EVOKE_B200_FOLDED_SYNC=0 run 29612 r2_t5_n2_barriers.json
EVOKE_B200_MASK_FREE=0 run 29613 r2_t5_n2_mask.json
EVOKE_B200_SCATTER_ROTATE=0 run 29614 r2_t5_n2_norotate.json
tail -c 600 gpurun_out/r2_t5_bench.err
