# round-2 final validation of the LAST build (collider clamp / tagged-wait slow path out of line) on ONE GPU: full GPU test
# suite, build() + smoke(), default bench, batch bench, then -- each after the same command exited 0 without a profiler --
# the ncu launch list and `ncu --set full` captures of tile_frame_kernel (fast / exact) and batch_frame_kernel (fast / exact)
set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -4 gpurun_out/r2_smoke.log | cut -c 1-200
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --workload batch4096 > gpurun_out/r2_bench_batch.json 2> gpurun_out/r2_bench_batch.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_batch.json')); print('batch', d['value'], d['roofline']['frac'], {k:(v['value'], v['roofline_frac']) for k,v in d.get('alt',{}).items()})" || tail -5 gpurun_out/r2_bench_batch.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --no-extra --arith fast"
$CMD > gpurun_out/plain_fast.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
cap() { # name kernel-regex cmd...
  name=$1; rx=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o gpurun_out/r2_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  python tools/ncu_summary.py gpurun_out/r2_$name.ncu-rep > gpurun_out/r2_${name}_summary.txt 2>&1
}
cap tile_fast tile_frame python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --no-extra --arith fast
cap tile_exact tile_frame python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --no-extra --arith exact
cap batch_fast batch_frame python bench.py --workload batch4096 --steps 1 --warmup 3 --no-cpu-baseline --arith fast
cap batch_exact batch_frame python bench.py --workload batch4096 --steps 1 --warmup 3 --no-cpu-baseline --arith exact
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
