t=${1:-r3n2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-budget 0 --cfg4 0 > gpurun_out/${t}_bench_n2.json 2> gpurun_out/${t}_bench_n2.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/${t}_bench_n2.json').read().strip().splitlines()[-1])
print('n2', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), d['parity'] and d['parity']['ok'], d['config'].get('ghost_exchange'))
for k,v in d['other_configs'].items(): print(k, {a:b for a,b in v.items() if a not in ('what','config','inner_cg_its')})
PY
tail -3 gpurun_out/${t}_bench_n2.err
