#!/bin/bash
# Stress of the runsimulation caller pool from C++ threads (examples/batch_main.cpp N T): every series must equal the
# one-batch run, for pools smaller than, equal to and larger than a warp, repeatedly.
set -e
cd "$(dirname "$0")/.."
g++ -std=c++17 -O2 -pthread -I include examples/batch_main.cpp -o /tmp/rs_pool_stress -L roadsurf_b200 -lroadsurf_b200 -Wl,-rpath,$PWD/roadsurf_b200
fail=0
for rep in 1 2 3 4 5; do
  for cfg in "97 3" "200 16" "401 64" "64 7" "333 33"; do
    set -- $cfg
    out=$(/tmp/rs_pool_stress $1 $2 | grep "series differing") || { echo "FAILED run: $cfg"; fail=1; continue; }
    echo "$cfg: ${out##*runsimulation from}"
    case "$out" in *"series differing: 0") ;; *) fail=1 ;; esac
  done
done
echo "pool stress: fail=$fail"
exit $fail
