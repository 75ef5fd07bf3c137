# One GPU round of evidence: tools/gr.sh 2400 1 'bash tools/gpu_round.sh r2u'   (tag = prefix of the files written under gpurun_out/)
t=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${t}_pytest.log 2>&1; tail -3 gpurun_out/${t}_pytest.log
python bench.py --impl reference --cpu-budget 10 > gpurun_out/${t}_ref.json 2> gpurun_out/${t}_ref.err
python bench.py > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; tail -c 600 gpurun_out/${t}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${t}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget 0 --cfg4 0 --parity 0 > gpurun_out/${t}_launch_bench.log 2>&1
ncu --set full --clock-control none -k regex:'ellip_kernel' -c 9 -o /tmp/${t}_sweep python tools/prof_kernels.py lin2 lin3 lin4 lin5 lin6 lin7 lin8 lin9 lin10 > gpurun_out/${t}_ncu_sweep.log 2>&1
ncu -i /tmp/${t}_sweep.ncu-rep --page raw --csv > gpurun_out/${t}_ncu_full_ellip_sweep.csv
ncu --set full --clock-control none --import-source on -k regex:'ellip_kernel|vec_kernel|q1_ellip_kernel|q1_elast_kernel' -c 12 -o /tmp/${t}_other python tools/prof_kernels.py q1 plap elast q1elast stokes > gpurun_out/${t}_ncu_other.log 2>&1
ncu -i /tmp/${t}_other.ncu-rep --page raw --csv > gpurun_out/${t}_ncu_full_q1_plap_elast_stokes.csv
cp /tmp/${t}_other.ncu-rep gpurun_out/ || true
cat gpurun_out/${t}_bench.json | head -c 3000
