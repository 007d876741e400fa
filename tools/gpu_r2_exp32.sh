# exp32 (last GPU seconds of the round): edge_delta_fast with the shortened chain (edge.so) vs the final build (final.so)
run() {
  PBD_B200_LIB=$PWD/tools/ab/$1 timeout 60 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith fast > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); r=d['roofline']; print('[$1 fast]', round(d['value'],1), round(r['frac'],4))" || tail -3 gpurun_out/ab.err
}
run final.so
run edge.so
run final.so
run edge.so
PBD_B200_LIB=$PWD/tools/ab/edge.so timeout 60 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fast and not p3 and not full_size" 2>&1 | tail -2
