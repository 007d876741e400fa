# final validation of the round's LAST build (exp30 + exp31 choices): A/B against exp30's build on this box first, then
# tools/gpu_r2_final2.sh (all GPU tests, smoke, default + batch bench, launch list, four ncu --set full captures)
run() {
  PBD_B200_LIB=$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith fast > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$2 fast]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
run $PWD/tools/ab/both.so exp30
run $PWD/cs121-softbodysim_b200/libpbd_b200.so final
bash tools/gpu_r2_final2.sh
