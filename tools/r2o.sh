mkdir -p gpurun_out
for v in $(ls dohp_b200/variants/); do
DOHP_B200_LIB=dohp_b200/variants/$v python tools/sweep_bench.py lin,jac 2,3,4,5,6,7,8,9,10 > gpurun_out/r2o_$v.json 2>gpurun_out/r2o_$v.err
cat gpurun_out/r2o_$v.json
done
