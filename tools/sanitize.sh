# compute-sanitizer over the parity tests: tools/gr.sh 900 1 'bash tools/sanitize.sh <tag> "<tools>" <ids file> <seconds per tool>'
# defaults: memcheck + racecheck (shared-memory phases) over one small test per kernel family (tools/sanitizer_ids.txt);
# tools/sanitizer_ids_all_degrees.txt holds every degree's instantiation of the fused kernels.  Logs: gpurun_out/<tag>_<tool>.txt
t=${1:-r3s}; tools=${2:-"memcheck racecheck"}; idf=${3:-tools/sanitizer_ids.txt}; lim=${4:-240}
set -f
ids=$(tr '\n' ' ' < $idf)
mkdir -p gpurun_out
for tool in $tools; do
  timeout $lim compute-sanitizer --tool $tool --print-limit 40 --log-file gpurun_out/${t}_$tool.txt python -m pytest $ids -q -p no:cacheprovider > gpurun_out/${t}_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/${t}_${tool}_pytest.log; grep -E "SUMMARY" gpurun_out/${t}_$tool.txt; head -c 1500 gpurun_out/${t}_$tool.txt
done
