for rep in 1 2; do for ln in 1 2 4; do
  timeout 300 python bench.py --lanes $ln --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('lanes $ln rep $rep', round(d['value'],1))"
done; done
for wl in config1 config2; do for ln in 1 4; do
  timeout 300 python bench.py --workload $wl --lanes $ln --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$wl lanes $ln', round(d['value'],1))"
done; done
