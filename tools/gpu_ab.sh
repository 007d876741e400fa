# usage: gpu_ab.sh libA libB ...   (paths relative to repo) -- alternates the builds on ONE box
for rep in 1 2; do
  for lib in "$@"; do
    PBD_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$lib rep $rep', round(d['value'],1), round(d['roofline']['frac'],4))"
  done
done
