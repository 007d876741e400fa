N=$1
PORT=29560
run() { name=$1; shift; PORT=$((PORT+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  python -c "import json; d=json.load(open('gpurun_out/bench_n${N}_$name.json')); print('N=$N $name', round(d['value'],1), d['unit'], d['scaling'], 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['sane'])" || (grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n${N}_$name.err | tail -5)
}
run weak --steps 5 --warmup 3 --no-cpu-baseline
run batch --workload batch4096 --steps 3 --warmup 2 --no-cpu-baseline
if [ "$2" = "shard" ]; then run shard8m --shard --workload big8m --steps 3 --warmup 3 --no-cpu-baseline; fi
