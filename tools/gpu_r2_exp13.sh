# exp13: how many independent sweep chains per SM?  2 x 256 threads (default) vs 3 x 160 / 4 x 128 / 5 x 96 / 3 x 128
# (all within 128 registers x 512 threads per SM); fast arithmetic, tagged hand-over, ONE box, alternating.
run() {
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --no-sustained --arith fast "$@" > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('[$*]', round(d['value'],1), round(d['roofline']['frac'],4), d['config'].get('schedule'))" || tail -3 gpurun_out/ab.err
}
for rep in 1 2; do
  run
  run --tiles-per-sm 4 --block-threads 128
  run --tiles-per-sm 3 --block-threads 160
  run --tiles-per-sm 3 --block-threads 128
  run --tiles-per-sm 5 --block-threads 96
done
