for rep in 1 2; do
  for lib in tools/ab/mixed.so tools/ab/pf.so; do
    PBD_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$lib rep $rep', round(d['value'],1), round(d['roofline']['frac'],4))"
  done
done
PBD_B200_LIB=$PWD/tools/ab/pf_trace.so PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/tr.json 2> gpurun_out/tr.err
grep "pbd-" gpurun_out/tr.err | tail -16 | cut -c1-400
