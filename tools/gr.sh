#!/bin/bash
# gpurun with retries on "no slot right now" (exit 3): tools/gr.sh <timeout> <gpus> '<command>'
t=$1; g=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $t -- "$@"; else /usr/local/graft/bin/gpurun --gpus $g --timeout $t -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
