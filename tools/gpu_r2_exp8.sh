set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "normals or collider" > gpurun_out/r2_exp8_pytest.log 2>&1; tail -3 gpurun_out/r2_exp8_pytest.log
run() { # env flags
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained --arith fast $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$1 [$2]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'steps', d['schedule']['edge_colors'], 'grid', d['schedule']['grid_blocks'], d['schedule']['block_threads'], 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
run "A=1" ""
for mt in 1024 512 256 128 64; do
  run "PBD_PLAN_MINTILE=$mt" "--workload config2"
  run "PBD_PLAN_MINTILE=$mt" "--workload config2 --block-threads 256"
done
for mt in 1024 512 256 128; do
  run "PBD_PLAN_MINTILE=$mt" "--workload config1"
  run "PBD_PLAN_MINTILE=$mt" "--workload config1 --block-threads 256"
done
run "PBD_PLAN_MINTILE=256" "--workload config2 --arith exact"
run "PBD_PLAN_MINTILE=256" "--workload small"
run "A=1" "--workload small"
