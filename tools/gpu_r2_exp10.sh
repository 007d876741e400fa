set -x
run() { # env flags
  env $1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sustained --arith fast $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$1 [$2]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'tiles', d['schedule']['tiles'], 'steps', d['schedule']['edge_colors'], 'grid', d['schedule']['grid_blocks'], d['schedule']['block_threads'], 'plan_ms', round(d['plan_ms']), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
run "A=1" "--workload big8m"
run "A=1" "--workload big8m --plan-sms 296"
run "A=1" "--workload big8m --plan-sms 592"
run "A=1" "--workload big8m --plan-sms 1184"
run "A=1" "--workload big8m --tiles-per-sm 1"
PBD_TILE_TRACE=1 timeout 300 python bench.py --workload config2 --steps 3 --warmup 3 --no-cpu-baseline --no-sustained --arith fast > gpurun_out/q.json 2> gpurun_out/r2_exp10_trace_config2.err; grep "pbd-" gpurun_out/r2_exp10_trace_config2.err | grep -v steps | tail -8
