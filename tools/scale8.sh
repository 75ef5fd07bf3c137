mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2w_bench_n8.json 2> gpurun_out/r2w_bench_n8.err
tail -c 400 gpurun_out/r2w_bench_n8.err; head -c 1500 gpurun_out/r2w_bench_n8.json
python tools/pcie_probe.py > gpurun_out/r2w_pcie_probe_n8.txt 2>&1; tail -3 gpurun_out/r2w_pcie_probe_n8.txt
