N=$1
set -x
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --workload big32m --steps 3 --warmup 3 --no-cpu-baseline --no-sustained --arith fast > gpurun_out/r2_big32m_n1.json 2> gpurun_out/r2_big32m_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --shard --workload big32m --steps 3 --warmup 3 --no-cpu-baseline --no-sustained --arith fast > gpurun_out/r2_big32m_n$N.json 2> gpurun_out/r2_big32m_n$N.err
fi
python -c "import json; d=json.load(open('gpurun_out/r2_big32m_n$N.json')); print('N=$N big32m', d['config']['backend'], round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'tiles', d['schedule']['tiles'], 'grid', d['schedule']['grid_blocks'], 'plan_ms', round(d['plan_ms']), 'init_ms', round(d['init_ms']), 'sane', d['sane'])" || tail -5 gpurun_out/r2_big32m_n$N.err
