timeout 600 python -m pytest tests -m gpu -x -q -k "tile" > gpurun_out/pytest_tile.log 2>&1; tail -3 gpurun_out/pytest_tile.log
for cfg in "interleaved 4" "interleaved 1"; do set -- $cfg
  PBD_TILE_TRACE=1 timeout 300 python bench.py --backend tile --order $1 --lanes $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v4_$1_$2.json 2> gpurun_out/bench_v4_$1_$2.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v4_$1_$2.json')); print('$1 lanes $2', d['value'], d['roofline']['frac'], d['schedule'])"
  grep pbd- gpurun_out/bench_v4_$1_$2.err | tail -12
done
