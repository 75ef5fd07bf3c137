set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'q1_elast_kernel' -c 1 -o /tmp/r2j_q1elast python tools/q1_bench.py elast 6:24 > gpurun_out/r2j_ncu.log 2>&1
ncu -i /tmp/r2j_q1elast.ncu-rep --page raw --csv > gpurun_out/r2j_ncu_full_q1_elast.csv
ncu -i /tmp/r2j_q1elast.ncu-rep --page source --csv > gpurun_out/r2j_ncu_src_q1_elast.csv
tail -3 gpurun_out/r2j_ncu.log
