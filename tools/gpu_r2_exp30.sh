# exp30: (thr) the mixed table's fourth word = first tet thread instead of the tet count (no per-step rebuild of blockDim-1-tid on
# the tet warps), (chain) tet_delta_fast with a shorter dependent path (wS order, w_k * (-C/6) formed under the MUFU.RCP, no
# alpha * lambda when the multiplier is inert); ni.so = the committed build, both.so = thr + chain
PBD_B200_LIB=$PWD/tools/ab/both.so timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "interleaved or fast or tagged or kuhn26 or riding" 2>&1 | tail -2
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 $3 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2 $3]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run ni.so fast
  run both.so fast
  run thr.so fast
  run chain.so fast
  run ni.so exact
  run thr.so exact
done
run ni.so fast "--workload batch4096"
run both.so fast "--workload batch4096"
