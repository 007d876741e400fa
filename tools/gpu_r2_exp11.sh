set -x
run() { # env flags
  env $1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained --arith fast $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$1 [$2]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'steps', d['schedule']['edge_colors'], 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
run "A=1" ""
run "A=1" "--order riding"
run "PBD_PLAN_RIDERS=1" "--order riding"
run "A=1" ""
run "PBD_PLAN_RIDERS=1" "--order riding"
