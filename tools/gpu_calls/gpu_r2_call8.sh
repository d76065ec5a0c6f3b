#!/bin/bash
# 2 GPUs: ring with the sender-moves protocol (C harness + bench.py both arms), all-pairs, conformance subset across NVLink
NP=2
export ACX_WATCHDOG_MS=20000
O=gpurun_out/call8; mkdir -p $O
R=baseline/_ref/bin
run() { timeout 120 ./bin/acxrun -np $NP --timeout 100 "$@" 2>&1 | tail -n 2; }
echo "### 0. N=1 ring"; for v in "X=1" "ACX_LATE_PUSH_MIN=0"; do echo "[$v] $(env $v ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | cut -c90-170)"; done
echo "### 1. conformance across NVLink"
for t in late_push tags host_reuse status_errors large_msgs stress ordering ring_all_graph graph_large; do echo "$t: $(run build/tests/$t | tail -n 1)"; done
ACX_FAULT_DELAY_US=200 run build/tests/late_push | tail -n 1
echo "### 2. ring 256 MiB, C harness"
for v in "X=1" "ACX_LATE_PUSH_MIN=0" "ACX_FAULT_DELAY_US=100" "ACX_LATE_PUSH_MIN=0 ACX_FAULT_DELAY_US=100"; do echo "[$v] $(env $v ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np $NP --timeout 100 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | cut -c90-190)"; done | tee $O/ring.txt
echo "[reference] $(run $R/p2p_bench ring --bytes 268435456 --no-heap | tail -n 1 | cut -c90-190)" | tee -a $O/ring.txt
echo "### 3. bench.py N=2 both arms (full)"
for impl in reference acx; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $NP --impl $impl > $O/bench_${impl}_n2.json 2> $O/bench_${impl}_n2.err
  echo "bench.py impl=$impl rc=$?"; grep '^{' $O/bench_${impl}_n2.json | tail -n 1 | cut -c1-2600; tail -n 2 $O/bench_${impl}_n2.err
done
echo "### done"
