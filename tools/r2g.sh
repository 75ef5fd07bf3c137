set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
python bench.py --impl reference --cpu-budget 10 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget 0 --cfg4 0 --parity 0 > gpurun_out/r2g_launch_bench.log 2>&1
ncu --set full --clock-control none -k regex:'ellip_kernel' -c 9 -o /tmp/r2g_sweep python tools/prof_kernels.py lin2 lin3 lin4 lin5 lin6 lin7 lin8 lin9 lin10 > gpurun_out/r2g_ncu_sweep.log 2>&1
ncu -i /tmp/r2g_sweep.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_full_ellip_sweep.csv
ncu --set full --clock-control none --import-source on -k regex:'ellip_kernel|vec_kernel|q1_ellip_kernel|q1_elast_kernel' -c 12 -o /tmp/r2g_other python tools/prof_kernels.py q1 plap elast stokes > gpurun_out/r2g_ncu_other.log 2>&1
ncu -i /tmp/r2g_other.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_full_q1_plap_elast_stokes.csv
ls -la /tmp/*.ncu-rep
cp /tmp/r2g_other.ncu-rep gpurun_out/ || true
tail -3 gpurun_out/r2g_pytest.log
