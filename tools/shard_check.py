"""torchrun --nproc-per-node N tools/shard_check.py : the multi-process (CUDA IPC) path of a sharded body.
Rank r drives GPU r; all ranks step the same body; positions must equal a single-GPU run of the same
schedule (done by rank 0 afterwards) bit for bit."""
import importlib, os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("cs121-softbodysim_b200")
capi, mg = pkg.capi, pkg.meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
x0, tets, edges = mg.kuhn_grid(n)
prm = capi.SolverParams.default(substeps=5)
sms = torch.cuda.get_device_properties(local).multi_processor_count
opt = capi.Options(order_mode=capi.ORDER_INTERLEAVED)
sb = capi.ShardedBody(prm, x0, edges, tets, rank, world, dist, device=local, options=opt)
for _ in range(8):
    sb.step(1 / 60)
pos = sb.read_positions()
if rank == 0:
    ref = capi.Body(prm, x0, edges, tets, device=local, options=capi.Options(backend=capi.BACKEND_TILE, order_mode=capi.ORDER_INTERLEAVED, plan_sms=sms * world))
    ref.step_async(1 / 60, 8); ref.sync()
    want = ref.read_positions()
    print(f"shard_check n={n} world={world}: V={len(x0)} T={len(tets)} tiles={sb.info()['tiles']} bit-exact={np.array_equal(pos, want)} max|d|={np.abs(pos-want).max():.3e} min_y={pos[:,1].min():.4f}", flush=True)
dist.barrier()
dist.destroy_process_group()
