# final build at 2 GPUs, launched as the driver launches it: headline (weak) + batch4096 (strong) + the sharded 8.4M-tet body
set -x
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_final_bench_n$N.json'))
    print('N=$N headline', round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'alt', {k:round(v['value'],1) for k,v in d.get('alt',{}).items()})
    for k,v in d.get('splits',{}).items():
        print('  split', k, {q: (round(v[q],2) if isinstance(v.get(q),(int,float)) else v.get(q)) for q in ('value','ms_per_step','error','sane')}, 'frac', v.get('roofline',{}).get('frac'), 'check', v.get('bit_identity_check'), 'alt', {a:round(b['value'],1) for a,b in v.get('alt',{}).items()} if isinstance(v.get('alt'),dict) else None)
except Exception as e:
    print('parse failed', e)
PY
tail -5 gpurun_out/r2_final_bench_n$N.err
