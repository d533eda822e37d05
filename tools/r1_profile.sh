#!/bin/bash
# Round-1 final evidence on ONE B200: full GPU suite, the default bench (with cpu baseline), the reference arm,
# the ncu launch list and one full ncu capture of the hot kernels.  Outputs under gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "reference arm rc=$?"
kill $SMI
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-clocks"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'tc_kernel|w_scale_kernel' -s 12 -c 4 -o gpurun_out/final_prof $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
for f in ("final_bench_n1", "final_bench_ref"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],4), "ms", round(d["value"]/1e6,4), "Mpairs/s", d.get("clocks"), d.get("cpu_baseline"), "e2e", d["e2e"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
