for bt in 512 256; do
  PBD_TILE_TRACE=1 timeout 300 python bench.py --backend tile --order interleaved --lanes 1 --block-threads $bt --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v5_$bt.json 2> gpurun_out/bench_v5_$bt.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v5_$bt.json')); print('bt $bt', d['value'], d['roofline']['frac'], d['schedule'])"
  grep pbd- gpurun_out/bench_v5_$bt.err | tail -8
done
