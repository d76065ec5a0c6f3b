#!/bin/bash
# 4 GPUs: the copy-engine rendezvous with more than one hop of host drift — conformance, ring 256 MiB (C harness), bench.py N=4.
NP=4
export ACX_WATCHDOG_MS=20000
O=gpurun_out/call14; mkdir -p $O
R=baseline/_ref/bin
ms() { tail -n 1 | grep -o '"ms_per_step": [0-9.e-]*\|"gbs_per_gpu": [0-9.e-]*' | tr '\n' ' '; }
echo "### 1. conformance at np=4"
echo "[ce_path 8 MiB, CE from 4 MiB] $(ACX_CE_MIN=4194304 timeout 120 ./bin/acxrun -np $NP --timeout 100 build/tests/ce_path 2>&1 | tail -n 2 | tr '\n' ' ')"
echo "[ce_path jitter] $(ACX_CE_MIN=4194304 ACX_FAULT_DELAY_US=200 timeout 120 ./bin/acxrun -np $NP --timeout 100 build/tests/ce_path 2>&1 | tail -n 2 | tr '\n' ' ')"
for t in late_push stress; do echo "$t: $(timeout 120 ./bin/acxrun -np $NP --timeout 100 build/tests/$t 2>&1 | tail -n 1)"; done
echo "### 2. ring 256 MiB, C harness, 40 steps"
for rep in 1 2; do
for v in "X=1" "ACX_FAULT_DELAY_US=100" "ACX_CE_MIN=0"; do echo "[$v] $(env $v ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np $NP --timeout 100 bin/p2p_bench ring --bytes 268435456 --steps 40 --warmup 5 2>&1 | ms)"; done
echo "[reference] $(timeout 120 ./bin/acxrun -np $NP --timeout 100 $R/p2p_bench ring --bytes 268435456 --steps 40 --warmup 5 --no-heap 2>&1 | ms)"
done | tee $O/ring.txt
echo "### 3. bench.py N=4 both arms (headline + e2e, no extras)"
for impl in reference acx; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus $NP --impl $impl --no-extras > $O/bench_${impl}_n4.json 2> $O/bench_${impl}_n4.err
  echo "bench.py impl=$impl rc=$?"; grep '^{' $O/bench_${impl}_n4.json | tail -n 1 | cut -c1-900
done
echo "### done"
