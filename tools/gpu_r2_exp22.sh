# exp22: home-tile slots sorted by their shifted tiles (fewer half-used sectors in the vertex loads / stores) vs caller order
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "resident or full_size or kuhn26" > gpurun_out/r2_exp22_pytest.log 2>&1; tail -3 gpurun_out/r2_exp22_pytest.log
set +x
run() {
  env $1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run PBD_PLAN_NOSIGSORT=1 fast
  run PBD_X=0 fast
  run PBD_PLAN_NOSIGSORT=1 exact
  run PBD_X=0 exact
done
