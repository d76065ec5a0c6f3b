#!/bin/bash
# One-GPU ncu captures of the three hot kernels (B200_PROFILING.md recipe). Reports land in gpurun_out/.
export ACX_WATCHDOG_MS=0 ACX_HEAP_BYTES=1g
O=gpurun_out; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
# launch list with device times (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_ring.csv bin/p2p_bench ring --bytes 268435456 --steps 4 --warmup 2 > /dev/null 2>&1
$NCU -k regex:mover -s 2 -c 2 -o $O/prof_mover -f bin/p2p_bench ring --bytes 268435456 --steps 4 --warmup 2 > $O/ncu_mover.log 2>&1; echo "mover rc=$?"
$NCU -k regex:halo_stencil_step -s 2 -c 2 -o $O/prof_halo -f bin/halo_bench fused --n 512 --steps 4 --warmup 2 --no-graph > $O/ncu_halo.log 2>&1; echo "halo rc=$?"
$NCU -k regex:gemm_fp8_kernel -s 1 -c 2 -o $O/prof_gemm -f build/tests/gemm_fp8_test 8192 8192 8192 2 > $O/ncu_gemm.log 2>&1; echo "gemm rc=$?"
ls -la $O/*.ncu-rep
