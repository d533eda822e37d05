#!/bin/bash
# Round-2 measurement recipes (run through gpurun; outputs land in gpurun_out/, the ones that matter are copied to profiles/).
#   gpurun            -- 'bash tools/r2_profile.sh n1'      full GPU suite, default bench + reference arm, config lines, launch list
#   gpurun            -- 'bash tools/r2_profile.sh k3'      K3 epilogue variants A/B (tests + bench)
#   gpurun --gpus N   -- 'bash tools/r2_profile.sh dist N'  multi-GPU parity tests of world N (and 2)
#   gpurun --gpus N   -- 'bash tools/r2_profile.sh ab N'    sharded bench with the A/B switches of profiles/r2_experiments.md §5
#   gpurun --gpus N   -- 'bash tools/r2_profile.sh bench N [extra bench.py flags]'
#   gpurun            -- 'bash tools/r2_profile.sh ncu TAG'  ncu --set full of K3 / K4t / K4b -> gpurun_out/TAG_prof_hot.ncu-rep
#   gpurun            -- 'bash tools/r2_profile.sh final'    n1 + the f1 line + smoke + ncu
set -u
mode=${1:-n1}
O=gpurun_out
mkdir -p $O
tr() { local port=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@"; }
case $mode in
n1)
  python -m pytest tests -m gpu -q > $O/r2_tests_n1.log 2>&1; echo "tests rc=$?" | tee -a $O/r2_tests_n1.log; tail -4 $O/r2_tests_n1.log
  python bench.py --steps 20 --warmup 5 > $O/r2_final_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
  python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_final_bench_ref.json 2>> $O/r2_bench_n1.err; echo "ref rc=$?"
  for c in cfg1 cfg2 cfg4; do python bench.py --config $c --steps 10 --warmup 3 > $O/r2_bench_${c}_n1.json 2>> $O/r2_bench_n1.err; echo "$c rc=$?"; done
  python bench.py --config cfg5 --steps 10 --warmup 3 > $O/r2_bench_cfg5_n1.json 2>> $O/r2_bench_n1.err; echo "cfg5 rc=$?"
  C="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cuda-eager --no-graph --no-clocks --no-dropin --no-sustained"
  $C > $O/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_eager_bench.csv $C > $O/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  ;;
k3)
  for v in 0 3 1; do EVK_K3_VARIANT=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager --no-dropin > $O/r2_bench_k3v$v.json 2>> $O/r2_bench_k3.err; echo "v$v rc=$?"; done
  EVK_K3_VARIANT=3 python -m pytest tests/test_gpu_estrip.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q > $O/r2_tests_k3v3.log 2>&1; tail -2 $O/r2_tests_k3v3.log
  ;;
dist)
  N=${2:-2}
  python -m pytest tests/test_gpu_distributed.py -m gpu -q > $O/r2_dist_tests_n$N.log 2>&1; echo "dist rc=$?" | tee -a $O/r2_dist_tests_n$N.log; tail -4 $O/r2_dist_tests_n$N.log
  ;;
ab)
  N=${2:-8}
  F="--steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager --no-kernel-events --no-dropin --no-sustained"
  tr 29621 $F > $O/r2_ab_n${N}_default.json 2>> $O/r2_ab.err
  EVOKE_B200_SIDE_BY_SIDE_CTAS=0 tr 29622 $F > $O/r2_ab_n${N}_sbs0.json 2>> $O/r2_ab.err
  EVOKE_B200_FOLDED_SYNC=0 tr 29623 $F > $O/r2_ab_n${N}_barriers.json 2>> $O/r2_ab.err
  EVOKE_B200_SCATTER_ROTATE=0 tr 29624 $F > $O/r2_ab_n${N}_norotate.json 2>> $O/r2_ab.err
  EVOKE_B200_MASK_FREE=0 tr 29625 $F > $O/r2_ab_n${N}_mask.json 2>> $O/r2_ab.err
  ;;
bench)
  N=${2:-8}; shift 2
  tr 29631 --steps 20 --warmup 5 --no-cpu-baseline --no-cuda-eager "$@" > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err; echo "bench rc=$?"
  ;;
ncu)
  T=${2:-r2}
  C="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-eager --no-graph --no-clocks --no-dropin --no-sustained --no-kernel-events"
  $C > $O/${T}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|w_scale" -s 12 -c 4 -o $O/${T}_prof_hot $C > $O/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  ;;
final)
  bash "$0" n1
  python bench.py --config f1 --steps 50 --warmup 5 > $O/r2_bench_f1_n1.json 2>> $O/r2_bench_n1.err; echo "f1 rc=$?"
  python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2_smoke.log
  bash "$0" ncu r2d
  ;;
esac
