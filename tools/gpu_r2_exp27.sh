# exp27: fast arithmetic: tets may take their vertices in any role order (pack_rows_relabel) vs PBD_PLAN_NORELABEL=1; tile + batch
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fast or batch" > gpurun_out/r2_exp27_pytest.log 2>&1; tail -3 gpurun_out/r2_exp27_pytest.log
set +x
run() {
  env $1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith fast $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4), d['schedule'].get('gather_wavefronts_permille'))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run PBD_PLAN_NORELABEL=1 ""
  run PBD_X=0 ""
done
run PBD_PLAN_NORELABEL=1 "--workload batch4096"
run PBD_X=0 "--workload batch4096"
