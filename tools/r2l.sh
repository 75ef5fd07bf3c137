set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_q1.py tests/test_gpu_elast_stokes.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1
tail -2 gpurun_out/r2l_pytest.log
(python tools/q1_bench.py elast 6:32 6:48 4:48
MORPH=0.2 python tools/q1_bench.py elast 6:32
for v in tj3 tj2; do echo $v; DOHP_B200_LIB=dohp_b200/variants/lib_$v.so python tools/q1_bench.py elast 6:32; done) > gpurun_out/r2l_q1_elast.txt 2>&1
cat gpurun_out/r2l_q1_elast.txt
