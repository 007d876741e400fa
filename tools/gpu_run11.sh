timeout 1200 python -m pytest tests -m gpu -x -q -k "tile or interleaved or batch or full_size" > gpurun_out/pytest_t.log 2>&1; tail -3 gpurun_out/pytest_t.log
for mode in flags grid; do
  if [ $mode = grid ]; then export PBD_TILE_GRIDSYNC=1; fi
  PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v13_$mode.json 2> gpurun_out/bench_v13_$mode.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v13_$mode.json')); print('$mode', d['value'], d['roofline']['frac'])"
  timeout 300 python bench.py --order strict --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v13s_$mode.json 2>/dev/null
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v13s_$mode.json')); print('$mode strict', d['value'], d['roofline']['frac'])"
done
grep "pbd-trace" gpurun_out/bench_v13_flags.err | tail -4 | cut -c 1-200
