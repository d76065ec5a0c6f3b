#!/bin/bash
# 8 GPUs: the headline at N=8 both arms, late protocol on/off, all-pairs (default / peer-branch / reference / NCCL), conformance
NP=8
export ACX_WATCHDOG_MS=20000
O=gpurun_out/call9; mkdir -p $O
R=baseline/_ref/bin
run() { timeout 150 ./bin/acxrun -np $NP --timeout 120 "$@" 2>&1 | tail -n 2; }
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "### 1. conformance at np=8"
for t in late_push tags stress ring_all_graph; do echo "$t: $(run build/tests/$t | tail -n 1)"; done
echo "### 2. ring 256 MiB, C harness"
for v in "X=1" "ACX_LATE_PUSH_MIN=0" "ACX_FAULT_DELAY_US=100" "ACX_LATE_PUSH_MIN=0 ACX_FAULT_DELAY_US=100"; do echo "[$v] $(env $v ACX_HEAP_BYTES=1g timeout 150 ./bin/acxrun -np $NP --timeout 120 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | cut -c90-190)"; done | tee $O/ring.txt
echo "[reference] $(run $R/p2p_bench ring --bytes 268435456 --no-heap | tail -n 1 | cut -c90-190)" | tee -a $O/ring.txt
echo "### 3. bench.py N=8: reference (full), acx (full), acx late off (no extras)"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
timeout 900 $T --master-port 29721 bench.py --gpus $NP --impl reference > $O/bench_reference_n8.json 2> $O/bench_reference_n8.err; echo "ref rc=$?"; grep '^{' $O/bench_reference_n8.json | tail -n 1 | cut -c1-2600
timeout 900 $T --master-port 29722 bench.py --gpus $NP > $O/bench_acx_n8.json 2> $O/bench_acx_n8.err; echo "acx rc=$?"; grep '^{' $O/bench_acx_n8.json | tail -n 1 | cut -c1-2600; tail -n 2 $O/bench_acx_n8.err | cut -c1-300
ACX_LATE_PUSH_MIN=0 timeout 600 $T --master-port 29723 bench.py --gpus $NP --no-extras > $O/bench_acx_n8_nolate.json 2>/dev/null; echo "acx late-off rc=$?"; grep '^{' $O/bench_acx_n8_nolate.json | tail -n 1 | cut -c1-700
echo "### 4. all-pairs graphs"
for b in 1048576 16777216; do
  for v in "X=1" "ACX_PEER_BRANCH=1" "ACX_LATE_PUSH_MIN=0"; do echo "allpairs bytes=$b [$v]: $(env $v ACX_HEAP_BYTES=2g timeout 150 ./bin/acxrun -np $NP --timeout 120 bin/p2p_bench allpairs --bytes $b 2>&1 | tail -n 1 | cut -c60-230)"; done
  echo "allpairs bytes=$b reference: $(run $R/p2p_bench allpairs --bytes $b --no-heap | tail -n 1 | cut -c60-230)"
  timeout 200 $T --master-port 29724 bench/nccl_p2p.py allpairs --bytes $b --graph 2>/dev/null | grep '^{' | cut -c1-230
done | tee $O/allpairs.txt
echo "### done"
