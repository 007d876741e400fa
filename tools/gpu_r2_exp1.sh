# round-2 experiment 1: tagged hand-over and fast arithmetic, parity first, then A/B on ONE box
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "tagged or fast_arith" > gpurun_out/r2_exp1_pytest.log 2>&1; tail -5 gpurun_out/r2_exp1_pytest.log
for flags in "" "--tagged" "--fast" "--fast --tagged" "" "--fast --tagged"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $flags > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('[$flags]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
done
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fast --tagged > gpurun_out/q.json 2> gpurun_out/r2_exp1_trace_fast_tagged.err; grep "pbd-" gpurun_out/r2_exp1_trace_fast_tagged.err | tail -12
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/q.json 2> gpurun_out/r2_exp1_trace_exact.err; grep "pbd-" gpurun_out/r2_exp1_trace_exact.err | tail -12
for wl in config2 config1; do for flags in "" "--fast --tagged"; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline $flags > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$wl [$flags]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
done; done
