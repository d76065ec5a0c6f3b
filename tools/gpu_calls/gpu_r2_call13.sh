#!/bin/bash
# 2 GPUs: the copy-engine ring step after the trace-driven changes — early release from wait kernels with few helpers (default),
# gate without an event (ACX_CE_SPINGATE), memOp completion / gate (ACX_CE_MEMOP) — C harness, then bench.py N=2 both arms.
NP=2
export ACX_WATCHDOG_MS=20000
O=gpurun_out/call13; mkdir -p $O
R=baseline/_ref/bin
ms() { tail -n 1 | grep -o '"ms_per_step": [0-9.e-]*'; }
echo "### 1. correctness of the variants across NVLink (ce_path, 8 MiB messages)"
for v in "X=1" "ACX_CE_SPINGATE=1" "ACX_CE_MEMOP=1" "ACX_CE_MEMOP=2" "ACX_CE_SPINGATE=1 ACX_CE_MEMOP=1"; do echo "[$v] $(env $v ACX_CE_MIN=4194304 timeout 120 ./bin/acxrun -np $NP --timeout 100 build/tests/ce_path 2>&1 | tail -n 1)"; done
echo "### 2. ring 256 MiB, C harness, 40 steps"
for rep in 1 2; do
for v in "X=1" "ACX_WAIT_EARLY_HELPERS=0" "ACX_CE_SPINGATE=1" "ACX_CE_MEMOP=1" "ACX_CE_MEMOP=2" "ACX_CE_SPINGATE=1 ACX_CE_MEMOP=1" "ACX_CE_MIN=0" "ACX_CE_MIN=0 ACX_WAIT_EARLY_HELPERS=0"; do echo "[$v] $(env $v ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np $NP --timeout 100 bin/p2p_bench ring --bytes 268435456 --steps 40 --warmup 5 2>&1 | ms)"; done
echo "[reference] $(timeout 120 ./bin/acxrun -np $NP --timeout 100 $R/p2p_bench ring --bytes 268435456 --steps 40 --warmup 5 --no-heap 2>&1 | ms)"
done | tee $O/ring.txt
echo "### 3. bench.py N=2 both arms (headline + e2e, no extras)"
for impl in reference acx; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus $NP --impl $impl --no-extras > $O/bench_${impl}_n2.json 2> $O/bench_${impl}_n2.err
  echo "bench.py impl=$impl rc=$?"; grep '^{' $O/bench_${impl}_n2.json | tail -n 1 | cut -c1-900
done
echo "### done"
