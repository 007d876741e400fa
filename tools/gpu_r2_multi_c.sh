N=$1
set -x
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --shard --steps 5 --warmup 3 --no-cpu-baseline --no-sustained "$@" > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('N=$N $name', d['config']['backend'], round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'tiles', d['schedule']['tiles'], 'grid', d['schedule']['grid_blocks'], 'plan_ms', round(d['plan_ms']), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
run big8m_tagged --workload big8m --arith fast
run big8m_tagged_exact --workload big8m --arith exact
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sustained --arith fast --workload big8m > gpurun_out/q.json 2> gpurun_out/q.err
python -c "import json; d=json.load(open('gpurun_out/q.json')); print('N=1 big8m', d['config']['backend'], round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'tiles', d['schedule']['tiles'], 'grid', d['schedule']['grid_blocks'], 'plan_ms', round(d['plan_ms']))" || tail -5 gpurun_out/q.err
