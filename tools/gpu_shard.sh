nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_shard_gpu.py -x -q 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/shard_check.py 20 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -5
