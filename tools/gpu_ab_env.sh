# usage: gpu_ab_env.sh "ENV1=.. ENV2=.." "ENVX=.." ...   -- alternates env settings on ONE box (default build)
for rep in 1 2; do
  for e in "$@"; do
    env $e timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('[$e] rep $rep', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab.err
  done
done
