# round-2 experiment 5: pipelined tagged epilogue (A/B against the previous commit on one box) + full tile parity
set -x
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_robustness_gpu.py -x -q -m gpu -k "not p3 and not full_size and not stream" > gpurun_out/r2_exp5_pytest.log 2>&1; tail -5 gpurun_out/r2_exp5_pytest.log
run() { # lib flags
  PBD_B200_LIB=$PWD/$1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $2 > gpurun_out/q.json 2> gpurun_out/q.err
  python -c "import json; d=json.load(open('gpurun_out/q.json')); print('$1 [$2]', d['config']['backend'], round(d['value'],1), round(d['roofline']['frac'],4), 'sane', d['sane'])" || tail -5 gpurun_out/q.err
}
M=cs121-softbodysim_b200/libpbd_b200.so
for rep in 1 2; do
  run tools/ab/prev.so "--tagged"; run $M "--tagged"; run tools/ab/prev.so "--fast --tagged"; run $M "--fast --tagged"; run tools/ab/prev.so ""; run $M ""
done
PBD_TILE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fast --tagged > gpurun_out/q.json 2> gpurun_out/r2_exp5_trace.err; grep "pbd-" gpurun_out/r2_exp5_trace.err | grep -v steps | tail -8
