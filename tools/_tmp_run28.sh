bash tools/r2_profile.sh n1
python bench.py --config f1 --steps 50 --warmup 5 > gpurun_out/r2_bench_f1_n1.json 2>> gpurun_out/r2_bench_n1.err; echo "f1 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
C="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-eager --no-graph --no-clocks --no-dropin --no-sustained --no-kernel-events"
$C > gpurun_out/r2d_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|w_scale" -s 12 -c 4 -o gpurun_out/r2d_prof_hot $C > gpurun_out/r2d_ncu_full.log 2>&1; echo "ncu full rc=$?"
