timeout 1200 python -m pytest tests -m gpu -x -q -k "tile or interleaved or batch" > gpurun_out/pytest_t.log 2>&1; tail -3 gpurun_out/pytest_t.log
PBD_TILE_TRACE=1 PBD_DUMP_TILE=gpurun_out/tile.bin timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v15.json 2> gpurun_out/bench_v15.err
python -c "import json,sys; d=json.load(open('gpurun_out/bench_v15.json')); print('v15', d['value'], d['roofline']['frac'])"
grep "pbd-" gpurun_out/bench_v15.err | tail -12 | grep -E "phase [1]" | cut -c 1-400
timeout 300 python bench.py --workload batch4096 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_batch3.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/bench_batch3.json')); print('batch', d['value'], d['roofline']['frac'])"
./tools/mb_sweep gpurun_out/tile.bin | head -4
