mkdir -p gpurun_out
python -m pytest tests/test_gpu_ellip.py tests/test_gpu_partition.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
python tools/sweep_bench.py lin,jac > gpurun_out/r2q_pipe1.json 2> gpurun_out/r2q_pipe1.err; cat gpurun_out/r2q_pipe1.json; tail -3 gpurun_out/r2q_pipe1.err
DOHP_B200_LIB=dohp_b200/variants/lib_pipe0.so python tools/sweep_bench.py lin,jac > gpurun_out/r2q_pipe0.json 2> gpurun_out/r2q_pipe0.err; cat gpurun_out/r2q_pipe0.json
MORPH=0.3 python tools/sweep_bench.py lin,jac > gpurun_out/r2q_pipe1_morph.json 2> gpurun_out/r2q_pipe1_morph.err; cat gpurun_out/r2q_pipe1_morph.json
MORPH=0.3 DOHP_B200_LIB=dohp_b200/variants/lib_pipe0.so python tools/sweep_bench.py lin,jac > gpurun_out/r2q_pipe0_morph.json 2> gpurun_out/r2q_pipe0_morph.err; cat gpurun_out/r2q_pipe0_morph.json
