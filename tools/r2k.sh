set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_q1.py tests/test_gpu_elast_stokes.py tests/test_gpu_partition.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
tail -4 gpurun_out/r2k_pytest.log
python tools/q1_bench.py elast 6:32 6:48 4:48 2:48 8:24 > gpurun_out/r2k_q1_elast.txt 2>&1
python tools/q1_bench.py ellip 4:64 2:64 6:40 8:32 > gpurun_out/r2k_q1_ellip.txt 2>&1
cat gpurun_out/r2k_q1_elast.txt gpurun_out/r2k_q1_ellip.txt
