mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/r2z_mgpu2.log 2>&1; echo rc=$?
tail -6 gpurun_out/r2z_mgpu2.log
for v in 1; do
DOHP_B200_P2P=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-budget 0 --cfg4 0 > gpurun_out/r2z_bench_n2_p2p$v.json 2> gpurun_out/r2z_bench_n2_p2p$v.err; echo rc=$?
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2z_bench_n2_p2p$v.json').read().strip().splitlines()[-1])
    print('p2p=$v', round(d['value'],1), [e['ms'] for e in d['roofline']['per_p']], d['parity'] and d['parity']['ok'], round(d['e2e']['value'],2))
except Exception as e: print('fail', e)
PY
done
tail -3 gpurun_out/r2z_bench_n2_p2p1.err
