set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c 1-150
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.json
timeout 600 python bench.py --order strict --no-cpu-baseline > gpurun_out/bench_strict.json 2> gpurun_out/bench_strict.err
timeout 600 python bench.py --workload batch4096 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tile_frame -s 1 -c 1 -o gpurun_out/prof_tile_r1e -f $CMD > gpurun_out/ncu_full.log 2>&1
CMDB="python bench.py --workload batch4096 --steps 1 --warmup 1 --no-cpu-baseline"
$CMDB > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:batch_frame -s 1 -c 1 -o gpurun_out/prof_batch_r1 -f $CMDB > gpurun_out/ncu_batch.log 2>&1
tail -2 gpurun_out/ncu_batch.log
