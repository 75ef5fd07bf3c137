mkdir -p gpurun_out
python -m pytest tests/test_gpu_q1.py -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; tail -12 gpurun_out/r2x_pytest.log
