mkdir -p gpurun_out
python -m pytest tests/test_gpu_lowdim.py tests/test_gpu_efs.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; tail -15 gpurun_out/r2s_pytest.log
