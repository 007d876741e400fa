timeout 900 python -m pytest tests -m gpu -x -q -k "tile or interleaved" > gpurun_out/pytest_t.log 2>&1; tail -3 gpurun_out/pytest_t.log
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err
python -c "import json,sys; d=json.load(open('gpurun_out/bench_v10.json')); print('v10', d['value'], d['roofline']['frac'], d['init_ms'])"
grep "pbd-" gpurun_out/bench_v10.err | tail -12 | grep -E "phase [1]" | cut -c 1-400
