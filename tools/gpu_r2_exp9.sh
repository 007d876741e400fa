set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "batch" > gpurun_out/r2_exp9_pytest.log 2>&1; tail -3 gpurun_out/r2_exp9_pytest.log
runb() { # env flags
  env $1 timeout 600 python bench.py --workload batch4096 --steps 5 --warmup 3 --no-cpu-baseline $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('batch $1 [$2]', round(d['value'],1), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), 'alt', {k: round(v['value'],1) for k,v in d.get('alt',{}).items()}, 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
runb "A=1" ""
runb "PBD_BATCH_NOSPLIT=1" ""
run() { # flags
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sustained $1 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('[$1]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'alt', {k: round(v['value'],1) for k,v in d.get('alt',{}).items()}, 'grid', d['schedule']['grid_blocks'], d['schedule']['block_threads'], 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
run "--workload config2"
run "--workload config1"
run "--workload small"
run ""
