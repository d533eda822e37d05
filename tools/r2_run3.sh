python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_distributed.py > gpurun_out/r2_t3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_t3_tests.log; tail -5 gpurun_out/r2_t3_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > gpurun_out/r2_t3_bench_maskfree.json 2> gpurun_out/r2_t3_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_t3_bench.err
EVOKE_B200_MASK_FREE=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager > gpurun_out/r2_t3_bench_mask.json 2>> gpurun_out/r2_t3_bench.err; echo "bench2 rc=$?"
python bench.py --steps 20 --warmup 5 --config cfg1 > gpurun_out/r2_t3_bench_cfg1.json 2>> gpurun_out/r2_t3_bench.err; echo "cfg1 rc=$?"
python bench.py --steps 20 --warmup 5 --config cfg2 --no-cpu-baseline > gpurun_out/r2_t3_bench_cfg2.json 2>> gpurun_out/r2_t3_bench.err; echo "cfg2 rc=$?"
python bench.py --steps 10 --warmup 3 --config cfg4 --no-cpu-baseline > gpurun_out/r2_t3_bench_cfg4.json 2>> gpurun_out/r2_t3_bench.err; echo "cfg4 rc=$?"
tail -c 800 gpurun_out/r2_t3_bench.err
