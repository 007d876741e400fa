# exp31: more of the tet warps' dependent path.  both.so = the committed build (exp30); all.so = tet test first + ride switch per
# visit + PTX index unpack (LOP3/SHF + LEA) + 1/36 folded + n_k w_k t products formed under the MUFU.RCP; ef.so = all but edge
# test first, idxc.so = all but the C index unpack, nopm.so = all but the pre-multiplied products
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 $3 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2 $3]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run both.so fast
  run all.so fast
  run ef.so fast
  run idxc.so fast
  run nopm.so fast
done
run both.so exact
run all.so exact
PBD_B200_LIB=$PWD/tools/ab/all.so timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fast or riding" 2>&1 | tail -2
