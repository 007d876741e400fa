[ -f gpurun_out/tile.bin ] || PBD_DUMP_TILE=gpurun_out/tile.bin timeout 300 python bench.py --backend tile --order interleaved --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
./tools/mb_sweep gpurun_out/tile.bin
