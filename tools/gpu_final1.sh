for wl in config2 config1 small; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), 'cpu', round(d['cpu_baseline']['value'],2), d['schedule'])"
done
timeout 300 python bench.py --workload config2 --backend stream --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config2_stream.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/bench_config2_stream.json')); print('config2 stream', round(d['value'],1), round(d['roofline']['frac'],4))"
timeout 300 python bench.py --backend stream --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stream_final.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/bench_stream_final.json')); print('headline stream', round(d['value'],1), round(d['roofline']['frac'],4))"
