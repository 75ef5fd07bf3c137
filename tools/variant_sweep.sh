mkdir -p gpurun_out
what=${1:-stokes}; export MORPH=${3:-0}
python tools/sweep_bench.py $what $2 > gpurun_out/r2v_base.json 2>gpurun_out/r2v_base.err; cat gpurun_out/r2v_base.json
for v in $(ls dohp_b200/variants/); do
DOHP_B200_LIB=dohp_b200/variants/$v python tools/sweep_bench.py $what $2 > gpurun_out/r2v_$v.json 2>gpurun_out/r2v_$v.err
cat gpurun_out/r2v_$v.json
done
