# exp28: resident visits skip the generic->async proxy fence when no lambda store follows (cf.so) vs always (base.so)
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith fast > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2 3; do
  run base.so
  run cf.so
done
