N=$1
set -x
timeout 900 python -m pytest tests/test_shard_gpu.py tests/test_robustness_gpu.py -x -q -m gpu > gpurun_out/r2_shard_pytest.log 2>&1; tail -4 gpurun_out/r2_shard_pytest.log
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --shard --steps 5 --warmup 3 --no-cpu-baseline --no-sustained "$@" > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('N=$N $name', d['config']['backend'], round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
run big8m_tagged --workload big8m --arith fast
run big8m_tagged_exact --workload big8m --arith exact
