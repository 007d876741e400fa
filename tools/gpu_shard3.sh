PORT=29530
run() { name=$1; shift; PORT=$((PORT+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "import json; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', round(d['value'],1), d['scaling'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['schedule']['tiles'], d['schedule']['grid_blocks'], d['sane'])" || (grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_$name.err | tail -5)
}
run shard2_headline --shard --steps 5 --warmup 3 --no-cpu-baseline
run shard2_big8m --shard --workload big8m --steps 3 --warmup 3 --no-cpu-baseline
run weak2_headline --steps 5 --warmup 3 --no-cpu-baseline
