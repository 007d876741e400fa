# round-2 experiment 2: raw shared addresses, record prefetch, store-phase trace
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "tagged or fast_arith or (interleaved_order_bit_exact and kuhn8) or strict_lanes or set_params" > gpurun_out/r2_exp2_pytest.log 2>&1; tail -5 gpurun_out/r2_exp2_pytest.log
run() { # lib flags
  PBD_B200_LIB=$PWD/$1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$1 [$2]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
M=cs121-softbodysim_b200/libpbd_b200.so
for rep in 1 2; do
  run $M ""; run tools/ab/pf.so ""; run $M "--fast"; run tools/ab/pf.so "--fast"
done
run $M "--fast --tiles-per-sm 2"
run tools/ab/pf.so "--fast --tiles-per-sm 2"
run $M "--fast --partitions 3"
run $M "--fast --partitions 5"
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fast > gpurun_out/q.json 2> gpurun_out/r2_exp2_trace_fast.err; grep "pbd-" gpurun_out/r2_exp2_trace_fast.err | grep -v steps | tail -8
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fast --tagged > gpurun_out/q.json 2> gpurun_out/r2_exp2_trace_fast_tagged.err; grep "pbd-" gpurun_out/r2_exp2_trace_fast_tagged.err | grep -v steps | tail -8
