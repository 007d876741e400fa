# exp21: four tagged vertex loads in flight per thread (tiles of 769..1024 vertices load in one round trip) vs three
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run cur.so fast
  run l4.so fast
done
run cur.so exact
run l4.so exact
