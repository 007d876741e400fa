set -x
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_wire_gpu.py tests/test_robustness_gpu.py -x -q -m gpu -k "collider or wire or server or robust or wrap or rank or unlaunched" > gpurun_out/r2_exp7_pytest.log 2>&1; tail -5 gpurun_out/r2_exp7_pytest.log
run() { # env flags
  env $1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$1 [$2]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'steps', d['schedule']['edge_colors'], 'grid', d['schedule']['grid_blocks'], d['schedule']['block_threads'], 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
for rep in 1 2; do
run "A=1" "--arith fast"
run "A=1" "--arith fast --tiles-per-sm 2"
run "A=1" "--arith exact"
run "A=1" "--arith exact --tiles-per-sm 2"
done
run "A=1" "--arith fast --tiles-per-sm 3"
run "A=1" "--arith fast --tiles-per-sm 3 --block-threads 160"
run "A=1" "--arith fast --tiles-per-sm 2 --block-threads 320"
run "A=1" "--arith fast --tiles-per-sm 2 --block-threads 384"
run "A=1" "--arith fast --tiles-per-sm 2 --order riding"
