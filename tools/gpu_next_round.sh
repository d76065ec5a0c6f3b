#!/bin/bash
# First GPU pass of the NEXT round: validates and measures everything that was written after round 1's GPU budget was
# spent (none of it is on a default path except bench.py's pre-posted order for >= 3 GPUs and the self-transfer grid).
#   gpurun --gpus 2 --timeout 900 -- tools/gpu_next_round.sh 2
#   gpurun --gpus 8 --timeout 900 -- tools/gpu_next_round.sh 8
NP=${1:-2}
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
R=baseline/_ref/bin
run() { timeout 120 ./bin/acxrun -np $NP --timeout 100 "$@" 2>&1 | tail -n 2; }

echo "### 1. new conformance program (PROC_NULL, communicator contexts, wildcard rejection)"
for np in 1 $NP; do timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/procnull_comms 2>&1 | tail -n 3; done

for np in 1 2 $NP; do timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/datatypes 2>&1 | tail -n 3; done
for np in 2 4; do timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/subcomm 2>&1 | tail -n 3; done

echo "### 1b. experimental device-side wildcards (ACX_WILDCARDS=1)"
for np in 1 2 $NP; do ACX_WILDCARDS=1 timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/wildcards 2>&1 | tail -n 3; done
timeout 60 ./bin/acxrun -np 2 --timeout 50 build/tests/wildcards 2>&1 | tail -n 1   # default: rejected cleanly

echo "### 1c. arrival-order jitter (ACX_FAULT_DELAY_US=200): the second-arriver logic with both orders on hardware"
for t in stress large_msgs ordering graph_large; do ACX_FAULT_DELAY_US=200 run build/tests/$t | tail -n 1; done

echo "### 2. examples"
make examples > /dev/null 2>&1
for e in stream_ring partitioned_producer; do timeout 60 ./bin/acxrun -np $NP --timeout 50 build/examples/$e 2>&1 | tail -n $NP; done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29701 examples/torch_ring.py 2>&1 | grep "rank" | head -n $NP

echo "### 3. ACX_PEER_BRANCH=1 (whole-operation per-peer branches): conformance, then all-pairs + ring against the default"
for t in ring ring_all ring_all_device ring_all_graph ring_graph_construction large_msgs ordering stress regcache; do
  ACX_PEER_BRANCH=1 run build/tests/$t | tail -n 1
done
for b in 1048576 16777216; do
  for br in 0 1; do
    echo "allpairs bytes=$b branch=$br: $(ACX_PEER_BRANCH=$br ACX_HEAP_BYTES=1g run bin/p2p_bench allpairs --bytes $b | tail -n 1)" | tee -a $O/allpairs_branch_n${NP}.txt
  done
  echo "allpairs bytes=$b reference: $(run $R/p2p_bench allpairs --bytes $b --no-heap | tail -n 1)" | tee -a $O/allpairs_branch_n${NP}.txt
done
for br in 0 1; do echo "ring branch=$br: $(ACX_PEER_BRANCH=$br ACX_HEAP_BYTES=1g run bin/p2p_bench ring --bytes 268435456 | tail -n 1)"; done
for br in 0 1; do echo "pingpong branch=$br: $(ACX_PEER_BRANCH=$br timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench pingpong 2>&1 | tail -n 1 | cut -c1-400)"; done

echo "### 3'. MPIX_Send_enqueue / MPIX_Recv_enqueue (eager receive = one fused post+wait kernel: 2 launches per ping-pong iteration instead of 3)"
for np in 1 $NP; do timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/sendrecv_enqueue 2>&1 | tail -n 2; done
for bl in "" "--blocking"; do echo "pingpong $bl: $(timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench pingpong --max 65536 $bl 2>&1 | tail -n 1 | cut -c1-500)"; done
for sr in "" "--sendrecv"; do for b in 8 4096; do echo "ring $b B $sr: $(run bin/p2p_bench ring --bytes $b --steps 200 --warmup 20 $sr | tail -n 1 | cut -c1-200)"; done; done
for bl in "" "--blocking"; do echo "pingpong_graph $bl: $(timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench pingpong_graph --max 65536 $bl 2>&1 | tail -n 1 | cut -c1-500)"; done

echo "### 3a. ACX_COOP=1 (cooperative moving with chunk stealing): conformance, then the ring in 'simple' order (no pre-posting) against the default"
for t in large_msgs ring_all_device ring_all_graph ordering stress regcache; do ACX_COOP=1 run build/tests/$t 16777216 | tail -n 1; done
for coop in 0 1; do echo "ring coop=$coop: $(ACX_COOP=$coop ACX_HEAP_BYTES=1g run bin/p2p_bench ring --bytes 268435456 | tail -n 1)"; done
for coop in 0 1; do
  ACX_COOP=$coop timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29704 bench.py --gpus $NP --order simple --no-extras 2>/dev/null | grep '^{' | tail -n 1 | cut -c1-260
done

echo "### 3b. ACX_PBATCH=1 (batched partitioned Start/Wait): conformance + API-level halo step against the default and the reference"
for np in 1 $NP; do ACX_PBATCH=1 timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/ring_partitioned 2>&1 | tail -n 1; done
for pbv in 0 1; do echo "halo api pbatch=$pbv: $(ACX_PBATCH=$pbv run bin/halo_bench api --n 512 --steps 10 | tail -n 1)"; done
echo "halo api reference: $(run $R/halo_bench api --n 512 --steps 10 | tail -n 1)"

echo "### 3b'. ACX_MEMOPS=1 (partitioned Start / Wait as one cuStreamBatchMemOp per group: value writes / GEQ waits on device flags)"
# (write-value to peer-mapped memory works on this box: profiles/hw_probe.md; the question is the cost of a batch)
for np in 1 $NP; do ACX_MEMOPS=1 timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/ring_partitioned 2>&1 | tail -n 2; done
for np in 2 4; do ACX_MEMOPS=1 timeout 90 ./bin/acxrun -np $np --timeout 80 build/tests/subcomm 2>&1 | tail -n 1; done
for mo in 0 1; do echo "halo api memops=$mo: $(ACX_MEMOPS=$mo run bin/halo_bench api --n 512 --steps 10 | tail -n 1)"; done

echo "### 3b''. graphs with rendezvous-size messages relaunched past the ring depth (new conformance program)"
for np in 1 $NP; do timeout 120 ./bin/acxrun -np $np --timeout 100 build/tests/graph_large 2>&1 | tail -n 2; done

echo "### 3c. halo kernel: deeper own-row prefetch (variants 7 = PD2, 8 = PD3) against the default (5): numerics, then 512^3 step time"
for v in 5 7 8; do
  ACX_HALO_VARIANT=$v timeout 120 python - <<'PY' 2>&1 | tail -n 1
import os, sys, torch
sys.path.insert(0, os.getcwd())
import mpi_acx_b200 as acx
from mpi_acx_b200.ops import HaloStencil, stencil_reference
acx.init(0)
op = HaloStencil(256, 32, 64, (1, 1, 1)); torch.manual_seed(3)
u0 = torch.randn(64, 32, 256, device="cuda").to(torch.bfloat16); op.field().copy_(u0); op.step(5); torch.cuda.synchronize()
print("variant", os.environ["ACX_HALO_VARIANT"], "max abs err", (op.field().float() - stencil_reference(u0, steps=5).float()).abs().max().item())
op.close(); acx.finalize()
PY
  echo "halo fused variant=$v: $(ACX_HALO_VARIANT=$v run bin/halo_bench fused --n 512 --steps 30 | tail -n 1)"
done

echo "### 4. ring with pre-posted receives (C harness) and bench.py at N=$NP (auto order = prepost from 3 GPUs up)"
ACX_HEAP_BYTES=1g run bin/p2p_bench ring --bytes 268435456 --prepost | tee $O/ring_acx_prepost_n${NP}.json
run $R/p2p_bench ring --bytes 268435456 --no-heap --prepost | tee $O/ring_ref_prepost_n${NP}.json
for order in auto simple; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus $NP --order $order > $O/bench_acx_n${NP}_$order.json 2> $O/bench_acx_n${NP}_$order.err
  echo "bench.py order=$order rc=$?"; grep '^{' $O/bench_acx_n${NP}_$order.json | tail -n 1 | cut -c1-600
done

echo "### 5. NCCL p2p yardstick"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29703 bench/nccl_p2p.py"
timeout 300 $T ring --bytes 268435456 2>/dev/null | grep '^{' | tee $O/nccl_ring_n${NP}.json
timeout 300 $T allpairs --bytes 16777216 --graph 2>/dev/null | grep '^{' | tee $O/nccl_allpairs_n${NP}.json
timeout 300 $T allpairs --bytes 1048576 --graph 2>/dev/null | grep '^{'
[ "$NP" = 2 ] && timeout 600 $T pingpong 2>/dev/null | grep '^{' | tee $O/nccl_pingpong.json | cut -c1-600

echo "### 6. self-transfer (N=1): mover on the caller's stream vs the side stream"
for ss in 1 0; do echo "side_streams=$ss: $(ACX_SIDE_STREAMS=$ss ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c100-200)"; done
echo "reference: $(timeout 60 ./bin/acxrun -np 1 --timeout 50 $R/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 --no-heap 2>&1 | tail -n 1 | cut -c60-160)"

echo "### 6b. self-transfer (N=1, the driver's 1-GPU headline; reference arm 0.1223 ms): mover flavour x block x grid, then where the step time goes"
# ncu saw the mover alone at 97 us of the 129-us step (profiles/mover_ncu.md): the sweep decides the self-transfer default,
# the trace (trigger / move-begin / move-end / wait timestamps from the device ring) shows what the other ~30 us are.
if [ "$NP" = 2 ]; then
  for mv in ldst ldstcs ldst256 tma; do for blk in 0 65536 131072 262144; do for ctas in 0 296 592 1184; do
    echo "mover=$mv block=$blk ctas=$ctas: $(ACX_MOVER=$mv ACX_MOVER_BLOCK=$blk ACX_MOVER_MAX_CTAS=$ctas ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c100-170)"
  done; done; done | tee $O/self_copy_sweep.txt
  for pdl in 1 0; do echo "pdl=$pdl: $(ACX_PDL=$pdl ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c100-170)"; done
  ACX_TRACE=1 ACX_TRACE_DIR=$O ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 6 --warmup 2 > /dev/null 2>&1
  python tools/trace_to_chrome.py $O/acx_trace_rank0.jsonl -o $O/self_ring_trace.json 2>&1 | tail -n 2
fi
