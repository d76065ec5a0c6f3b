#!/bin/bash
# 1 GPU: the full GPU suite on the current tree, halo v5 (producer warp) timing + ncu, GEMM check.
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call10; mkdir -p $O
echo "### 1. full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 | tee $O/pytest.txt
echo "### 2. halo v5"
for i in 1 2; do ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1; done
ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench both --n 512 --steps 20 2>&1 | tail -n 1
ACX_HEAP_BYTES=2g ACX_RANK=0 ACX_SIZE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo_stencil_tma -s 3 -c 1 -f -o $O/halo_tma_v5 bin/halo_bench fused --n 512 --steps 4 --warmup 2 --no-graph > $O/ncu_halo.log 2>&1; tail -n 2 $O/ncu_halo.log
echo "### 3. gemm 8192^3 check + timing"
timeout 200 build/tests/gemm_fp8_test 8192 8192 8192 2>&1 | tail -n 3
echo "### 4. smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
echo "### done"
