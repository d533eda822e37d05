C="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-eager --no-graph --no-clocks --no-dropin --no-sustained --no-kernel-events"
$C > gpurun_out/r2c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|w_scale" -s 12 -c 4 -o gpurun_out/r2c_prof_hot $C > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2c_prof_hot.ncu-rep
