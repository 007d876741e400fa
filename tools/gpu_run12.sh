for cfg in "1 512" "2 256" "2 384" "3 160"; do set -- $cfg
  timeout 300 python bench.py --tiles-per-sm $1 --block-threads $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v14_$1_$2.json 2> gpurun_out/bench_v14_$1_$2.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v14_$1_$2.json')); print('tps $1 bt $2', d['value'], d['roofline']['frac'], d['schedule']['grid_blocks'], d['schedule']['tet_colors'])"
done
