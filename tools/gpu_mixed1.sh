# mixed colour steps: parity of the interleaved order first, then A/B on one box
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "interleaved" 2>&1 | tail -3
for rep in 1 2; do
  for cfg in "base.so PBD_X=0" "mixed.so PBD_PLAN_MIXED=0" "mixed.so PBD_X=0"; do
    set -- $cfg
    env PBD_B200_LIB=$PWD/tools/ab/$1 $2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python -c "import json,sys; d=json.load(open('gpurun_out/ab.json')); print('[$cfg] rep $rep', round(d['value'],1), round(d['roofline']['frac'],4))" || tail -3 gpurun_out/ab.err
  done
done
