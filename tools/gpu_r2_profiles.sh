# round-2 profile captures (each after its plain command exited 0 without a profiler)
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --arith fast"
$CMD > gpurun_out/plain_fast.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
cap() { # name kernel-regex cmd...
  name=$1; rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/r2_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
cap tile_fast tile_frame python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --arith fast
cap tile_exact tile_frame python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --arith exact
cap batch_fast batch_frame python bench.py --workload batch4096 --steps 1 --warmup 3 --no-cpu-baseline --arith fast
cap batch_exact batch_frame python bench.py --workload batch4096 --steps 1 --warmup 3 --no-cpu-baseline --arith exact
