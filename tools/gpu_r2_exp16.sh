# exp16: group-table entry of the next step fetched before the barrier (-DPBD_SWEEP_GROUP_AHEAD) vs base, ONE box;
# then CTA 0's visit budget with per-step stamps (trace build)
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run base.so fast
  run ga.so fast
  run base.so exact
  run ga.so exact
done
PBD_B200_LIB=$PWD/tools/ab/trace.so PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-sustained --arith fast > gpurun_out/trace.json 2> gpurun_out/r2_exp16_trace.txt
grep -c . gpurun_out/r2_exp16_trace.txt
