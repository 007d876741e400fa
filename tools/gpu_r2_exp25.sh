# exp25: tagged write-back stores prepared in batches of 1 / 2 / 3 before they are issued (the kernel sits at its register limit:
# the batches of 2 and 3 spill 24..64 bytes), ONE box
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run cur.so fast
  run s1.so fast
  run s2.so fast
  run s3.so fast
done
run cur.so exact
run s1.so exact
run s3.so exact
