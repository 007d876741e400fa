timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for lib in tools/ab/tabu.so tools/ab/cap.so; do
  PBD_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$lib interleaved', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab.err
  PBD_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --order strict > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$lib strict', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab.err
  PBD_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload batch4096 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$lib batch4096', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab.err
done
