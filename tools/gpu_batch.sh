timeout 900 python -m pytest tests -m gpu -x -q -k "batch" > gpurun_out/pytest_batch.log 2>&1; tail -15 gpurun_out/pytest_batch.log
timeout 900 python bench.py --workload batch4096 --steps 3 --warmup 2 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err; tail -5 gpurun_out/bench_batch.err; cat gpurun_out/bench_batch.json | head -c 2500
