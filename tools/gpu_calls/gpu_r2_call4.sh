#!/bin/bash
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call4; mkdir -p $O
echo "### 1. status_errors: which launch faults?"
for v in "ACX_PDL=0" "CUDA_LAUNCH_BLOCKING=1" "ACX_SIDE_STREAMS=0" "ACX_EAGER_MAX=512" "ACX_NCHAN=16 ACX_TAGDIR=0"; do echo "[$v] $(env $v timeout 100 ./bin/acxrun -np 1 --timeout 90 build/tests/status_errors 2>&1 | tail -n 2 | tr '\n' ' ')"; done
ACX_RANK=0 ACX_SIZE=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 build/tests/status_errors > $O/memcheck_direct.txt 2>&1; grep -E "Invalid|at 0x|by thread|Address 0x|in /|wait_kernel|trigger|ERROR SUMMARY|PASS|FAIL" $O/memcheck_direct.txt | head -30
echo "### 2. halo v3"
timeout 600 python -m pytest tests -m gpu -q -k "halo or late_push or exact_tag" 2>&1 | tail -n 8
for i in 1 2; do ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1; done
ACX_HEAP_BYTES=2g ACX_RANK=0 ACX_SIZE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo_stencil_tma -s 3 -c 1 -f -o $O/halo_tma_v3 bin/halo_bench fused --n 512 --steps 4 --warmup 2 --no-graph > $O/ncu_halo.log 2>&1; tail -n 2 $O/ncu_halo.log
echo "### 3. self ring with the late protocol"
ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 2>&1 | tail -n 1
ACX_LATE_PUSH_MIN=0 ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 2>&1 | tail -n 1
echo "### done"
