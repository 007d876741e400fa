# round-2 validation 1: full GPU test suite, smoke, default bench, P3 report
set -x
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -8 gpurun_out/r2_smoke.log | cut -c 1-200
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 1500 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
timeout 900 python bench.py --workload batch4096 > gpurun_out/r2_bench_batch.json 2> gpurun_out/r2_bench_batch.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_batch.json')); print('batch', d['value'], d['roofline']['frac'], d.get('alt'))" || tail -5 gpurun_out/r2_bench_batch.err
timeout 600 python bench.py --workload batch4096 --lanes 2 --arith exact --no-cpu-baseline > gpurun_out/q.json 2> gpurun_out/q.err; python -c "
import json; d=json.load(open('gpurun_out/q.json')); print('batch lanes2 exact', d['value'], d['roofline']['frac'])" || tail -5 gpurun_out/q.err
timeout 900 python tools/p3_report.py > gpurun_out/p3_residuals.json 2> gpurun_out/p3_report.err; tail -8 gpurun_out/p3_report.err | cut -c 1-250
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 700 gpurun_out/r2_bench_reference.json; tail -3 gpurun_out/r2_bench_reference.err
