set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1
tail -5 gpurun_out/r2h_pytest.log
DOHP_B200_LIB=dohp_b200/variants/lib_ringf.so python tools/sweep_bench.py jac 2,3,4,5,6,7,8,9,10 > gpurun_out/r2h_ringf.json 2>gpurun_out/r2h_ringf.err
python tools/sweep_bench.py jac 2,3,4,5,6,7,8,9,10 > gpurun_out/r2h_base.json 2>gpurun_out/r2h_base.err
cat gpurun_out/r2h_ringf.json gpurun_out/r2h_base.json
