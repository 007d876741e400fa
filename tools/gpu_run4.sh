timeout 900 python -m pytest tests -m gpu -x -q -k "interleaved or strict_lanes" > gpurun_out/pytest_il.log 2>&1; tail -3 gpurun_out/pytest_il.log
for ln in 1 2 4; do
  PBD_TILE_TRACE=1 timeout 300 python bench.py --backend tile --order interleaved --lanes $ln --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6_$ln.json 2> gpurun_out/bench_v6_$ln.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v6_$ln.json')); print('lanes $ln', d['value'], d['roofline']['frac'])"
  grep "pbd-" gpurun_out/bench_v6_$ln.err | tail -12 | grep -E "phase [01]" 
done
