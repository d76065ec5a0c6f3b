#!/bin/bash
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call6; mkdir -p $O
echo "### 1. full GPU suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 | tee $O/pytest.txt
echo "### 2. halo v4"
for i in 1 2; do ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1; done
echo "### 3. self ring with the late protocol"
ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 2>&1 | tail -n 6
ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 16777216 --steps 20 --warmup 5 2>&1 | tail -n 3
ACX_LATE_PUSH_MIN=0 ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 2>&1 | tail -n 1
echo "### done"
