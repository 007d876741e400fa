# quick re-validation after a planner change: all GPU tests, then the three bench lines (no CPU baseline)
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for flags in "" "--order strict" "--workload batch4096 --steps 3"; do
  timeout 300 python bench.py --warmup 3 --no-cpu-baseline $flags > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('[$flags]', round(d['value'],1), round(d['roofline']['frac'],4), 'init_ms', round(d.get('init_ms',0)))" || tail -3 gpurun_out/q.err
done
