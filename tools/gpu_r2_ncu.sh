# ncu --set full captures of the frame kernel (one launch each), after the plain command exited 0
set -x
cap() { # name flags
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $2"
  $CMD > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:tile_frame -s 2 -c 1 -o gpurun_out/r2_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -2 gpurun_out/ncu_$1.log
}
cap fast "--fast"
cap riding_fast "--order riding --fast"
cap exact ""
