mkdir -p gpurun_out
python tools/sweep_bench.py elast,stokes > gpurun_out/r2r_base.json 2>gpurun_out/r2r_base.err; cat gpurun_out/r2r_base.json
for v in $(ls dohp_b200/variants/); do
DOHP_B200_LIB=dohp_b200/variants/$v python tools/sweep_bench.py elast,stokes > gpurun_out/r2r_$v.json 2>gpurun_out/r2r_$v.err
cat gpurun_out/r2r_$v.json
done
