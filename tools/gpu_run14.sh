PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err
python -c "import json,sys; d=json.load(open('gpurun_out/bench_v16.json')); print('v16', d['value'], d['roofline']['frac'])"
grep "pbd-" gpurun_out/bench_v16.err | tail -16 | grep -v trace | cut -c 1-330
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v16b.json 2>/dev/null
python -c "import json,sys; d=json.load(open('gpurun_out/bench_v16b.json')); print('v16 no trace', d['value'], d['roofline']['frac'])"
