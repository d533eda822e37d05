# 1 GPU: full GPU suite, K3 variant A/B (tests + bench), cfg4/cfg5 lines, launch list
python -m pytest tests -m gpu -q --deselect tests/test_gpu_distributed.py > gpurun_out/r2_t7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_t7_tests.log; tail -6 gpurun_out/r2_t7_tests.log
EVK_K3_VARIANT=3 python -m pytest tests/test_gpu_estrip.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2_t7_tests_v3.log 2>&1; echo "tests v3 rc=$?" >> gpurun_out/r2_t7_tests_v3.log; tail -4 gpurun_out/r2_t7_tests_v3.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager --no-dropin"
python bench.py $B > gpurun_out/r2_t7_bench_v0.json 2> gpurun_out/r2_t7_bench.err; echo "v0 rc=$?"
EVK_K3_VARIANT=3 python bench.py $B > gpurun_out/r2_t7_bench_v3.json 2>> gpurun_out/r2_t7_bench.err; echo "v3 rc=$?"
EVK_K3_VARIANT=1 python bench.py $B > gpurun_out/r2_t7_bench_v1.json 2>> gpurun_out/r2_t7_bench.err; echo "v1 rc=$?"
EVK_K3_VARIANT=3 python bench.py $B --config cfg4 --steps 10 > gpurun_out/r2_t7_bench_cfg4_v3.json 2>> gpurun_out/r2_t7_bench.err; echo "cfg4 v3 rc=$?"
python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_t7_bench_cfg5.json 2>> gpurun_out/r2_t7_bench.err; echo "cfg5 rc=$?"
tail -c 1200 gpurun_out/r2_t7_bench.err
