for st in 0 300 450 600; do
  PBD_TILE_STAGGER=$st timeout 300 python bench.py --tiles-per-sm 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v11_$st.json 2> gpurun_out/bench_v11_$st.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v11_$st.json')); print('tps2 stagger $st', d['value'], d['roofline']['frac'], d['schedule']['grid_blocks'])"
done
