for rep in 1 2; do
  for flags in "--tiles-per-sm 1" "--tiles-per-sm 2" ; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $flags > gpurun_out/ab.json 2> gpurun_out/ab.err
    python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('[$flags] rep $rep', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab.err
  done
done
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --tiles-per-sm 2 > gpurun_out/tr.json 2> gpurun_out/tr.err
grep "pbd-" gpurun_out/tr.err | grep -v steps | tail -8 | cut -c1-330
