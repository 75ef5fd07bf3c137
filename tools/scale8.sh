mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --cpu-budget 0 --cfg4 0 > gpurun_out/r2y_bench_n8_$tag.json 2> gpurun_out/r2y_bench_n8_$tag.err; echo rc=$?; python - <<PY
import json
d=json.loads(open('gpurun_out/r2y_bench_n8_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value'],1), [e['ms'] for e in d['roofline']['per_p']], d['parity'] and d['parity']['ok'], round(d['e2e']['value'],2))
PY
}
run p2p X=1
tail -3 gpurun_out/r2y_bench_n8_p2p.err
