# exp29: collide_vertex out of line (ni.so: the frame kernel shrinks from 6.7k to 3.7k instructions) vs inlined (base.so)
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 $2]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run base.so fast
  run ni.so fast
  run base.so exact
  run ni.so exact
done
PBD_B200_LIB=$PWD/tools/ab/ni.so timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "collider" 2>&1 | tail -2
