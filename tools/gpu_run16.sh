timeout 1200 python -m pytest tests -m gpu -x -q -k "tile or interleaved or batch or full_size" > gpurun_out/pytest_t.log 2>&1; tail -3 gpurun_out/pytest_t.log
bash tools/gpu_ab_env.sh "PBD_TILE_NO_EARLY_STORE=1" "PBD_X=0"
