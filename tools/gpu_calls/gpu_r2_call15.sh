#!/bin/bash
# 1 GPU: halo kernel — 3-D tensor boxes vs 1-D bulk copies of the same (contiguous) boxes, one vs two CTAs per SM, ncu of the default.
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call15; mkdir -p $O
hb() { env "$@" ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1 | grep -o '"ms_per_step": [0-9.e-]*'; }
echo "### 1. halo fused 512^3, ms per step"
for rep in 1 2; do
for v in "X=1" "ACX_HALO_BULK=1" "ACX_HALO_CTAS_PER_SM=1" "ACX_HALO_BULK=1 ACX_HALO_CTAS_PER_SM=1"; do echo "[$v] $(hb $v)"; done
done | tee $O/halo.txt
echo "### 2. numerics with the bulk loads"
ACX_HALO_BULK=1 timeout 600 python -m pytest tests -m gpu -q -k "halo or stencil" 2>&1 | tail -n 3
echo "### 3. ncu of the default kernel and of the bulk variant"
ACX_HEAP_BYTES=2g ACX_RANK=0 ACX_SIZE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:halo_stencil_tma -s 3 -c 1 -f -o $O/halo_v5 bin/halo_bench fused --n 512 --steps 4 --warmup 2 --no-graph > $O/ncu_halo.log 2>&1; tail -n 2 $O/ncu_halo.log
ACX_HALO_BULK=1 ACX_HEAP_BYTES=2g ACX_RANK=0 ACX_SIZE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:halo_stencil_tma -s 3 -c 1 -f -o $O/halo_v5_bulk bin/halo_bench fused --n 512 --steps 4 --warmup 2 --no-graph > $O/ncu_halo_bulk.log 2>&1; tail -n 2 $O/ncu_halo_bulk.log
echo "### done"
