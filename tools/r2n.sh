mkdir -p gpurun_out
DOHP_B200_LIB=dohp_b200/variants/lib_tpb128.so python tools/sweep_bench.py lin,jac 2,3,4,5,6,7,8,9,10 > gpurun_out/r2n_tpb128.json 2>gpurun_out/r2n_tpb128.err
python tools/sweep_bench.py lin,jac 2,3,4,5,6,7,8,9,10 > gpurun_out/r2n_base.json 2>gpurun_out/r2n_base.err
cat gpurun_out/r2n_tpb128.json gpurun_out/r2n_base.json; tail -3 gpurun_out/r2n_tpb128.err
