set -x
mkdir -p gpurun_out
export DOHP_B200_LIB=dohp_b200/variants/lib_prof.so
for d in 0 1 4 2 8 16 6 14 30; do echo "DBG $d"; DOHP_Q1_DBG=$d python tools/q1_bench.py elast 6:32; done > gpurun_out/r2i_elast_ablation.txt 2>&1
for d in 0 1 4 2 6; do echo "DBG $d"; DOHP_Q1_DBG=$d python tools/q1_bench.py ellip 4:64; done > gpurun_out/r2i_ellip_ablation.txt 2>&1
cat gpurun_out/r2i_elast_ablation.txt gpurun_out/r2i_ellip_ablation.txt | grep -v "^+"
