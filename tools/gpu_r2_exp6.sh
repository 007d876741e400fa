# round-2 experiment 6: planner knobs with the tagged fast kernel; tile sizes for the small configs
set -x
run() { # env flags
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained --arith fast $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$1 [$2]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'steps', d['schedule']['edge_colors'], 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
run "A=1" ""
run "PBD_PLAN_CAPM=0" ""
run "PBD_PLAN_CAPM=2" ""
run "A=1" "--tiles-per-sm 2"
run "A=1" "--tiles-per-sm 2 --block-threads 512"
run "A=1" "--arith exact"
run "PBD_PLAN_CAPM=0" "--arith exact"
for tv in 0 128 256 400 600; do run "A=1" "--workload config2 --tile-vertices $tv"; done
for tv in 0 100 200 400 800; do run "A=1" "--workload config1 --tile-vertices $tv"; done
run "A=1" "--workload config1 --order riding"
run "A=1" "--workload config2 --order riding --tile-vertices 256"
run "A=1" "--workload big8m"
