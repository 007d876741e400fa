timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "interleaved" 2>&1 | tail -3
for rep in 1 2; do
  for lib in tools/ab/relB.so tools/ab/tabu.so; do
    PBD_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$lib rep $rep', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab.err
  done
done
