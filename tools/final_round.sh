t=${1:-r2final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${t}_pytest.log 2>&1; tail -3 gpurun_out/${t}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
python bench.py --impl reference --cpu-budget 10 > gpurun_out/${t}_ref.json 2> gpurun_out/${t}_ref.err
python bench.py > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; tail -c 300 gpurun_out/${t}_bench.err; head -c 600 gpurun_out/${t}_bench.json
