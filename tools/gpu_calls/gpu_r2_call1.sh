#!/bin/bash
# Round 2, GPU call 1 (2 GPUs): full GPU suite without -x, bandwidth probe, and the measurements that decide defaults.
NP=2
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call1; mkdir -p $O
R=baseline/_ref/bin
run() { timeout 120 ./bin/acxrun -np $NP --timeout 100 "$@" 2>&1 | tail -n 2; }
nvidia-smi topo -m > $O/topo.txt 2>&1; nvidia-smi -L | head -3; lscpu | grep -i "numa\|^CPU(s)\|Model name" ; numactl -H 2>/dev/null | head -5
echo "### A. pytest -m gpu (no -x)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 25 | tee $O/pytest.txt
echo "### B. bwprobe: ring 2 GPUs, uni, self"
top() { awk '{for(i=1;i<=NF;i++) if($i=="GB/s") print $(i-1), $0}' $1 | sort -rn | head -${2:-10} | cut -d' ' -f2-; }
timeout 200 tools/bwprobe 2 > $O/bwprobe_ring2.txt 2>&1; grep -c GB/s $O/bwprobe_ring2.txt; grep "copy engine" $O/bwprobe_ring2.txt; top $O/bwprobe_ring2.txt 12
timeout 200 tools/bwprobe 2 268435456 uni > $O/bwprobe_uni2.txt 2>&1; grep -c GB/s $O/bwprobe_uni2.txt; grep "copy engine" $O/bwprobe_uni2.txt; top $O/bwprobe_uni2.txt 8
timeout 200 tools/bwprobe 1 > $O/bwprobe_self1.txt 2>&1; grep -c GB/s $O/bwprobe_self1.txt; grep "copy engine" $O/bwprobe_self1.txt; top $O/bwprobe_self1.txt 8
echo "### C. small messages: pingpong default vs --blocking; ring default vs --sendrecv; reference"
for bl in "" "--blocking"; do echo "pingpong $bl: $(timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench pingpong --max 65536 $bl 2>&1 | tail -n 1 | cut -c1-700)"; done | tee $O/pingpong_small.txt
for sr in "" "--sendrecv"; do for b in 8 4096; do echo "ring $b B $sr: $(run bin/p2p_bench ring --bytes $b --steps 200 --warmup 20 $sr | tail -n 1 | cut -c1-260)"; done; done | tee $O/ring_small.txt
for b in 8 4096; do echo "ring $b B reference: $(run $R/p2p_bench ring --bytes $b --steps 200 --warmup 20 --no-heap | tail -n 1 | cut -c1-260)"; done | tee -a $O/ring_small.txt
echo "### D. large ring, C harness: default / coop / sender-push / reference"
echo "ring default: $(ACX_HEAP_BYTES=1g run bin/p2p_bench ring --bytes 268435456 | tail -n 1)" | tee $O/ring_large.txt
for t in large_msgs ring_all_device ordering stress; do ACX_COOP=1 run build/tests/$t 16777216 | tail -n 1; done
echo "ring coop: $(ACX_COOP=1 ACX_HEAP_BYTES=1g run bin/p2p_bench ring --bytes 268435456 | tail -n 1)" | tee -a $O/ring_large.txt
echo "ring push: $(ACX_SENDER_PUSH=1 ACX_HEAP_BYTES=1g run bin/p2p_bench ring --bytes 268435456 | tail -n 1)" | tee -a $O/ring_large.txt
echo "ring reference: $(run $R/p2p_bench ring --bytes 268435456 --no-heap | tail -n 1)" | tee -a $O/ring_large.txt
echo "### E. bench.py N=2 (simple order both arms)"
for impl in acx reference; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus $NP --impl $impl > $O/bench_${impl}_n2.json 2> $O/bench_${impl}_n2.err
  echo "bench.py impl=$impl rc=$?"; grep '^{' $O/bench_${impl}_n2.json | tail -n 1 | cut -c1-1500
done
echo "### F. API-level partitioned: halo api default / pbatch / memops / reference"
for v in "" "ACX_PBATCH=1" "ACX_MEMOPS=1"; do echo "halo api [$v]: $(env $v timeout 120 ./bin/acxrun -np $NP --timeout 100 bin/halo_bench api --n 512 --steps 10 2>&1 | tail -n 1)"; done | tee $O/halo_api.txt
echo "halo api reference: $(run $R/halo_bench api --n 512 --steps 10 | tail -n 1)" | tee -a $O/halo_api.txt
echo "### G. all-pairs graphs N=2 (default / peer-branch / reference)"
for b in 1048576 16777216; do
  for br in 0 1; do echo "allpairs bytes=$b branch=$br: $(ACX_PEER_BRANCH=$br ACX_HEAP_BYTES=1g run bin/p2p_bench allpairs --bytes $b | tail -n 1)"; done
  echo "allpairs bytes=$b reference: $(run $R/p2p_bench allpairs --bytes $b --no-heap | tail -n 1)"
done | tee $O/allpairs.txt
echo "### H. NCCL yardstick N=2"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29703 bench/nccl_p2p.py"
timeout 200 $T ring --bytes 268435456 2>/dev/null | grep '^{' | tee $O/nccl_ring_n2.json
timeout 200 $T allpairs --bytes 16777216 --graph 2>/dev/null | grep '^{' | tee $O/nccl_allpairs16m_n2.json
timeout 200 $T allpairs --bytes 1048576 --graph 2>/dev/null | grep '^{' | tee $O/nccl_allpairs1m_n2.json
timeout 400 $T pingpong 2>/dev/null | grep '^{' | tee $O/nccl_pingpong.json | cut -c1-900
echo "### done"
