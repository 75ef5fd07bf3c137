mkdir -p gpurun_out
python -m pytest tests/test_gpu_rotation.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; tail -15 gpurun_out/r2t_pytest.log
