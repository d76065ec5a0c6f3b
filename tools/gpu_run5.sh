#!/bin/bash
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
echo "### pytest gpu"
timeout 700 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 15
echo "### ours api ppgemm np=2 on one GPU (debug)"
ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/ppgemm_bench api --steps 2 --m 1024 --d 1024 2>&1 | tail -n 8
echo "### ncu"
tools/ncu_capture.sh 2>&1 | tail -n 12
