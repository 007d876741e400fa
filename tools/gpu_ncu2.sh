CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_ncu2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tile_frame -s 1 -c 1 -o gpurun_out/prof_tile_r1d -f $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
