CMD="python bench.py --backend tile --order interleaved --lanes 1 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_ncu1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tile_frame -s 1 -c 1 -o gpurun_out/prof_tile_r1b -f $CMD > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
