# 2 GPUs: sharded-body parity tests, then the 8.4M-tet body over both GPUs (mixed colour steps)
timeout 600 python -m pytest tests/test_shard_gpu.py -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_shard_8m.sh 2
