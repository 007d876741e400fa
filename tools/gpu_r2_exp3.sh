# round-2 experiment 3: riding edges
set -x
timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "riding and not p3" > gpurun_out/r2_exp3_pytest.log 2>&1; tail -5 gpurun_out/r2_exp3_pytest.log
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $1 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('[$1]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
for rep in 1 2; do
  run ""; run "--fast"; run "--order riding"; run "--order riding --fast"; run "--order riding --fast --tagged"
  PBD_B200_LIB=$PWD/tools/ab/smemriders.so run "--order riding --fast"
done
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fast --order riding > gpurun_out/q.json 2> gpurun_out/r2_exp3_trace.err; grep "pbd-" gpurun_out/r2_exp3_trace.err | grep -v steps | tail -8
for wl in config2 config1; do run "--workload $wl --order riding --fast"; done
