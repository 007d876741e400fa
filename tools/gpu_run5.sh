for ln in 1 2; do
  PBD_TILE_TRACE=1 timeout 300 python bench.py --backend tile --order interleaved --lanes $ln --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v7_$ln.json 2> gpurun_out/bench_v7_$ln.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v7_$ln.json')); print('lanes $ln', d['value'], d['roofline']['frac'])"
  grep "pbd-" gpurun_out/bench_v7_$ln.err | tail -12 | grep -E "phase [1]" 
done
./tools/mb_sweep gpurun_out/tile.bin | head -8
