#!/bin/bash
# compute-sanitizer passes over the single-rank conformance programs (the reference has no sanitizer
# usage at all). memcheck + racecheck + synccheck; results under gpurun_out/sanitize/.
export ACX_WATCHDOG_MS=120000
O=gpurun_out/sanitize; mkdir -p $O
for tool in memcheck racecheck synccheck; do
  for t in ring ring_all_device ring_partitioned; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 build/tests/$t > $O/${tool}_$t.log 2>&1
    echo "$tool $t rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${tool}_$t.log | tail -n 1)"
  done
done | tee $O/summary.txt
