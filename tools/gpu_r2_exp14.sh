# exp14: shared-memory placement search (pbd_placement.cpp) on/off, ONE box, alternating; parity of the touched paths first
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "p1 or full_size or kuhn26 or interleaved or fast" > gpurun_out/r2_exp14_pytest.log 2>&1; tail -4 gpurun_out/r2_exp14_pytest.log
set +x
run() {
  env $1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4), {k:round(v,4) for k,v in r.items() if k.endswith('_frac')})" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run PBD_PLAN_PLACE=0 fast
  run PBD_PLAN_PLACE=1 fast
  run PBD_PLAN_PLACE=0 exact
  run PBD_PLAN_PLACE=1 exact
done
