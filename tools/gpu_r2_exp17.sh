# exp17 (timing only, results of x1/x3 are NOT valid): what do thread 0's TMA duties cost the visit?
#   x1 = no lambda write-back, x3 = x1 + the record prefetch issued by thread 128 (a warp that idles in most colour steps)
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run base.so fast
  run x1.so fast
  run x3.so fast
done
run base.so exact
run x3.so exact
