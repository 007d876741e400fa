# exp18: lambda write-back with plain 16-byte stores by all threads (v6) vs two bulk stores by thread 0 (base), ONE box;
# parity first (lambda arrays are part of the bit-exact comparison)
set -x
PBD_B200_LIB=$PWD/tools/ab/v6.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_robustness_gpu.py -x -q -m gpu -k "p1 or full_size or kuhn26 or interleaved or fast or wrap or unlaunched or shard" > gpurun_out/r2_exp18_pytest.log 2>&1; tail -4 gpurun_out/r2_exp18_pytest.log
set +x
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run base.so fast
  run v6.so fast
  run base.so exact
  run v6.so exact
done
