# exp23: one 512-thread tile per SM (148 tiles, records twice as large, resident blocks still fit) vs two 256-thread tiles per SM
run() {
  timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith fast "$@" > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$*]', round(d['value'],1), round(r['frac'],4), d['config'].get('schedule'))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run
  run --tiles-per-sm 1
done
PBD_TILE_TRACE=1 timeout 120 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --no-sustained --arith fast --tiles-per-sm 1 2>&1 | grep -E "resident|ftrace" | head -6 | cut -c 1-250
