# exp26: placement search on the batch backend (whole body in shared memory) vs PBD_BATCH_NOPLACE=1; batch parity tests first
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "batch" > gpurun_out/r2_exp26_pytest.log 2>&1; tail -3 gpurun_out/r2_exp26_pytest.log
set +x
run() {
  env $1 timeout 300 python bench.py --workload batch4096 --steps 4 --warmup 3 --no-cpu-baseline --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run PBD_BATCH_NOPLACE=1 fast
  run PBD_X=0 fast
  run PBD_BATCH_NOPLACE=1 exact
  run PBD_X=0 exact
done
