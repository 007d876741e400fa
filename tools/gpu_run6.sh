for cfg in "1 512" "2 256" "3 160" "2 512"; do set -- $cfg
  PBD_TILE_TRACE=1 timeout 300 python bench.py --tiles-per-sm $1 --block-threads $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v8_$1_$2.json 2> gpurun_out/bench_v8_$1_$2.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v8_$1_$2.json')); print('tps $1 bt $2', d['value'], d['roofline']['frac'], d['schedule'])"
  grep "pbd-" gpurun_out/bench_v8_$1_$2.err | tail -12 | grep -E "phase [1]" | cut -c 1-400
done
