mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/r2z_mgpu4.log 2>&1; echo rc=$?
grep -v "^\*\|OMP_NUM" gpurun_out/r2z_mgpu4.log | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 10 --warmup 3 --cpu-budget 0 --cfg4 0 > gpurun_out/r2z_bench_n4_p2p.json 2> gpurun_out/r2z_bench_n4_p2p.err; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/r2z_bench_n4_p2p.json').read().strip().splitlines()[-1])
print('n4', round(d['value'],1), [e['ms'] for e in d['roofline']['per_p']], d['parity'] and d['parity']['ok'], round(d['e2e']['value'],2), d['config'].get('ghost_exchange'))
PY
