#!/bin/bash
# 2 GPUs: copy-engine rendezvous (ACX_CE_MIN) — conformance across NVLink, ring 256 MiB C harness (CE on / off / reference),
# size sweep for the threshold, bench.py N=2 both arms (headline only), N=1 ring.
NP=2
export ACX_WATCHDOG_MS=20000
O=gpurun_out/call11; mkdir -p $O
R=baseline/_ref/bin
run() { timeout 120 ./bin/acxrun -np $NP --timeout 100 "$@" 2>&1 | tail -n 3; }
cut1() { tail -n 1 | cut -c1-400; }
echo "### 1. ce_path across NVLink"
for v in "X=1" "ACX_CE_MIN=0" "ACX_FAULT_DELAY_US=200"; do echo "[$v] $(env $v CE_PATH_VERBOSE=1 timeout 120 ./bin/acxrun -np $NP --timeout 100 build/tests/ce_path 2>&1 | tail -n 3 | tr '\n' ' ')"; done
echo "[np=1] $(timeout 120 ./bin/acxrun -np 1 --timeout 100 build/tests/ce_path 2>&1 | tail -n 2 | tr '\n' ' ')"
for t in late_push large_msgs stress status_errors; do echo "$t: $(run build/tests/$t | tail -n 1)"; done
echo "### 2. ring 256 MiB, C harness"
for v in "X=1" "ACX_CE_MIN=0" "ACX_FAULT_DELAY_US=100" "ACX_CE_SPIN_US=0"; do echo "[$v] $(env $v ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np $NP --timeout 100 bin/p2p_bench ring --bytes 268435456 2>&1 | cut1)"; done | tee $O/ring.txt
echo "[reference] $(run $R/p2p_bench ring --bytes 268435456 --no-heap | cut1)" | tee -a $O/ring.txt
echo "### 3. ring size sweep: CE from 1 MiB vs off"
for b in 1048576 2097152 4194304 8388608 16777216 67108864; do
  for v in "ACX_CE_MIN=1048576" "ACX_CE_MIN=0"; do echo "[$b $v] $(env $v ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np $NP --timeout 100 bin/p2p_bench ring --bytes $b --steps 50 --warmup 10 2>&1 | tail -n 1 | grep -o '"ms_per_step": [0-9.e-]*')"; done
  echo "[$b reference] $(run $R/p2p_bench ring --bytes $b --steps 50 --warmup 10 --no-heap | tail -n 1 | grep -o '"ms_per_step": [0-9.e-]*')"
done | tee $O/sweep.txt
echo "### 4. N=1 ring"
for v in "X=1" "ACX_CE_MIN=0"; do echo "[$v] $(env $v ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 2>&1 | cut1)"; done | tee $O/ring_n1.txt
echo "[reference] $(timeout 60 ./bin/acxrun -np 1 --timeout 50 $R/p2p_bench ring --bytes 268435456 --no-heap 2>&1 | cut1)" | tee -a $O/ring_n1.txt
echo "### 5. bench.py N=2 both arms (headline + e2e, no extras)"
for impl in reference acx; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $NP --impl $impl --no-extras > $O/bench_${impl}_n2.json 2> $O/bench_${impl}_n2.err
  echo "bench.py impl=$impl rc=$?"; grep '^{' $O/bench_${impl}_n2.json | tail -n 1 | cut -c1-1800; tail -n 2 $O/bench_${impl}_n2.err
done
echo "### 6. bench.py N=1 (acx, no extras)"
timeout 300 python bench.py --no-extras > $O/bench_acx_n1.json 2> $O/bench_acx_n1.err; echo "rc=$?"; grep '^{' $O/bench_acx_n1.json | tail -n 1 | cut -c1-1500
echo "### done"
