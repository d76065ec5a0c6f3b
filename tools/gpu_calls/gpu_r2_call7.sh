#!/bin/bash
export ACX_WATCHDOG_MS=3000
O=gpurun_out/call7; mkdir -p $O
B="bin/p2p_bench ring --steps 10 --warmup 3"
r() { echo "[$1] $(env $1 ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 $B --bytes $2 2>&1 | tail -n 2 | cut -c1-230 | tr '\n' ' ')"; }
for sz in 33554432 67108864 134217728 268435456; do r "X=1" $sz; done
r "ACX_PDL=0" 268435456
r "ACX_SIDE_STREAMS=0" 268435456
for h in 8 64 148 200; do r "ACX_LATE_HELPERS=$h" 268435456; done
r "ACX_MOVER_MAX_CTAS=148" 268435456
r "ACX_MOVER_MAX_CTAS=296" 268435456
echo "### wildcard tags + late_push + host_reuse"
timeout 600 python -m pytest tests -m gpu -q -k "exact_tag or late_push or host_reuse or truncation" 2>&1 | tail -n 5
echo "### done"
