mkdir -p gpurun_out
python -m pytest tests/test_gpu_ellip.py tests/test_gpu_solver_parity.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -2 gpurun_out/r2p_pytest.log
ncu --set full --clock-control none --import-source on -k regex:'ellip_kernel' -c 3 -o /tmp/r2p_hi python tools/prof_kernels.py lin10 lin8 lin4 > gpurun_out/r2p_ncu.log 2>&1
ncu -i /tmp/r2p_hi.ncu-rep --page raw --csv > gpurun_out/r2p_ncu_full_lin10_lin8_lin4.csv
for k in 0 1 2; do ncu -i /tmp/r2p_hi.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > gpurun_out/r2p_src_$k.csv; done
python tools/sweep_bench.py lin,jac > gpurun_out/r2p_sweep.json 2> gpurun_out/r2p_sweep.err; cat gpurun_out/r2p_sweep.json
