set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/r2m_mgpu2.log 2>&1
tail -5 gpurun_out/r2m_mgpu2.log
