# N GPUs (argument): side-by-side contractions A/B
N=$1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager --no-kernel-events --no-dropin --no-sustained > gpurun_out/$2 2>> gpurun_out/r2_t9_bench.err; echo "$2 rc=$?"; }
run 29641 r2_t9_n${N}_sbs74.json
EVOKE_B200_SIDE_BY_SIDE_CTAS=0 run 29642 r2_t9_n${N}_sbs0.json
tail -c 300 gpurun_out/r2_t9_bench.err
