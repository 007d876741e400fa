run() { # name, env...
  name=$1; shift
  env "$@" PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v12_$name.json 2> gpurun_out/bench_v12_$name.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_v12_$name.json')); print('$name', d['value'], d['roofline']['frac'], d['schedule']['tet_colors'], d['plan_ms'])"
  grep "pbd-trace" gpurun_out/bench_v12_$name.err | tail -4 | head -2 | cut -c 1-200
}
run nobal PBD_PLAN_NOBAL=1
run cap0 PBD_X=1
run cap1 PBD_PLAN_CAPM=1
run nobal_ns4 PBD_PLAN_NOBAL=1 PBD_PLAN_NOSNAP=4
run cap1_ns4 PBD_PLAN_CAPM=1 PBD_PLAN_NOSNAP=4
