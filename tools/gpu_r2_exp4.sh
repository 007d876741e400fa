# round-2 experiment 4: 128-bit tagged hand-over
set -x
timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "tagged or (riding and not p3) or fast_arith" > gpurun_out/r2_exp4_pytest.log 2>&1; tail -5 gpurun_out/r2_exp4_pytest.log
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $1 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('[$1]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
for rep in 1 2; do
  run ""; run "--tagged"; run "--fast"; run "--fast --tagged"; run "--order riding --fast --tagged"
done
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fast --tagged > gpurun_out/q.json 2> gpurun_out/r2_exp4_trace.err; grep "pbd-" gpurun_out/r2_exp4_trace.err | grep -v steps | tail -8
for wl in config2 config1; do run "--workload $wl --fast --tagged"; done
