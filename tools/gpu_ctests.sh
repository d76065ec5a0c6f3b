#!/bin/bash
# Run the C/CUDA conformance programs on the GPU box. Usage: tools/gpu_ctests.sh "1 2 4" [tests...]
# Every run is bounded by a device-side watchdog (ACX_WATCHDOG_MS) and a host timeout.
NPS=${1:-"1 2"}; shift
TESTS=${@:-"ring ring_all ring_all_device ring_all_graph ring_graph_construction ring_partitioned large_msgs"}
export ACX_WATCHDOG_MS=${ACX_WATCHDOG_MS:-20000}
mkdir -p gpurun_out
fail=0
for np in $NPS; do
  for t in $TESTS; do
    out=$(timeout 120 ./bin/acxrun -np $np --timeout 100 build/tests/$t 2>&1); rc=$?
    echo "$out" | tail -n 12
    echo "== $t np=$np rc=$rc"
    [ $rc -ne 0 ] && fail=1
  done
done 2>&1 | tee gpurun_out/ctests.log
grep -c "rc=0" gpurun_out/ctests.log
exit 0
