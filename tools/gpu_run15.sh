timeout 1200 python -m pytest tests -m gpu -x -q -k "tile or interleaved or batch" > gpurun_out/pytest_t.log 2>&1; tail -3 gpurun_out/pytest_t.log
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v19.json 2> gpurun_out/bench_v19.err
python -c "import json,sys; d=json.load(open('gpurun_out/bench_v19.json')); print('v17 trace', d['value'], d['roofline']['frac'])"
grep "pbd-" gpurun_out/bench_v19.err | tail -16 | grep -E "wait|ftrace" | cut -c 1-300 | head -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v19b.json 2>/dev/null
python -c "import json,sys; d=json.load(open('gpurun_out/bench_v19b.json')); print('v17', d['value'], d['roofline']['frac'])"
timeout 300 python bench.py --workload batch4096 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_batch4.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/bench_batch4.json')); print('batch', d['value'], d['roofline']['frac'])"
