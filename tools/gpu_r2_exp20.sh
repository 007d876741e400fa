# exp20: resident record blocks (RES kernel: visits 0 and 2 keep their block in shared memory for the frame) vs streaming all four
set -x
PBD_TILE_TRACE=1 timeout 120 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --no-sustained --arith fast 2>&1 | grep "resident"
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fast or full_size" > gpurun_out/r2_exp20_pytest_a.log 2>&1; tail -3 gpurun_out/r2_exp20_pytest_a.log
set +x
run() {
  env $1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2 3; do
  run PBD_TILE_NORESIDENT=1 fast
  run PBD_X=0 fast
done
