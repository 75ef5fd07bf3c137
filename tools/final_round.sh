# Final evidence of a round: GPU tests, smoke, reference arm, bench line, and the ncu launch list of the same bench command.
#   tools/gr.sh 1200 1 'bash tools/final_round.sh r3final'
t=${1:-r3final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${t}_pytest.log 2>&1; tail -3 gpurun_out/${t}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
python bench.py --impl reference --cpu-budget 10 > gpurun_out/${t}_ref.json 2> gpurun_out/${t}_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; tail -c 300 gpurun_out/${t}_bench.err; head -c 700 gpurun_out/${t}_bench.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${t}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget 0 --cfg4 0 --parity 0 > gpurun_out/${t}_launch_bench.log 2>&1; echo launches rc=$?
