# exp19: fast mode without the inert tet multipliers (alpha == 0) vs carrying them (PBD_TILE_KEEP_LAMBDA=1), vs the previous build
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fast or full_size or p3" > gpurun_out/r2_exp19_pytest.log 2>&1; tail -4 gpurun_out/r2_exp19_pytest.log
set +x
run() {
  env $1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run PBD_B200_LIB=$PWD/tools/ab/base.so fast
  run PBD_TILE_KEEP_LAMBDA=1 fast
  run PBD_X=0 fast
done
run PBD_B200_LIB=$PWD/tools/ab/base.so exact
run PBD_X=0 exact
