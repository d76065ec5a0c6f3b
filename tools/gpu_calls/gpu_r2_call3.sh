#!/bin/bash
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call3; mkdir -p $O
echo "### 1. host-register probe"; timeout 60 tools/hostreg_probe 2>&1 | tail -n 30
echo "### 2. status_errors verbose"; ACX_VERBOSE=1 timeout 100 ./bin/acxrun -np 1 --timeout 90 build/tests/status_errors 2>&1 | tail -n 12
echo "### 3. halo v2: numerics + timing"
timeout 600 python -m pytest tests -m gpu -q -k "halo" 2>&1 | tail -n 6
ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench both --n 512 --steps 20 2>&1 | tail -n 3
for c in 1 2; do echo "ctas/sm=$c: $(ACX_HALO_CTAS_PER_SM=$c ACX_HEAP_BYTES=2g timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1)"; done
ACX_HEAP_BYTES=2g ACX_RANK=0 ACX_SIZE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo_stencil_tma -s 3 -c 1 -f -o $O/halo_tma_v2 bin/halo_bench fused --n 512 --steps 4 --warmup 2 --no-graph > $O/ncu_halo.log 2>&1; tail -n 2 $O/ncu_halo.log
echo "### done"
