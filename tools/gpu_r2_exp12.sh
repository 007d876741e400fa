set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "jacobi or normals" > gpurun_out/r2_exp12_pytest.log 2>&1; tail -5 gpurun_out/r2_exp12_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default.json')); print(round(d['value'],1), round(d['roofline']['frac'],4), {k:round(v,4) for k,v in d['roofline'].items() if k.endswith('_frac')}, 'alt', round(d['alt']['exact']['value'],1), 'extra', {k:(round(v['value'],1), round(v['roofline_frac'],4)) if 'value' in v else v for k,v in d.get('extra',{}).items()})" || tail -5 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --backend stream --steps 3 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/q.json 2> gpurun_out/q.err; python -c "
import json; d=json.load(open('gpurun_out/q.json')); print('stream', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/q.err
