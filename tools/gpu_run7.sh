#!/bin/bash
# 2-GPU pass: latency knobs (PDL, fused rendezvous), movers, registration cache.
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
R=baseline/_ref/bin
summ='import sys,json
j=json.loads(sys.stdin.read()); print([(r["bytes"], round(r["oneway_us"],1), round(r["gbs"],1)) for r in j["rows"]])'
echo "### halo variants (512^3), np=1 then np=2"
for v in 2 4 5 6; do ACX_HALO_VARIANT=$v timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/halo_bench fused --n 512 --steps 20 2>&1 | tail -n 1 | cut -c1-250; done
for v in 5 6; do ACX_HALO_VARIANT=$v timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/halo_bench both --n 256 --steps 5 2>&1 | tail -n 1; done
for v in 4 5; do ACX_HALO_VARIANT=$v timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/halo_bench fused --n 512 --steps 20 2>&1 | tail -n 1 | cut -c1-250; done
ACX_HALO_VARIANT=5 timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/halo_bench both --n 256 --steps 5 2>&1 | tail -n 1
echo "### ring_partitioned np=8 on 2 GPUs (debug)"
timeout 120 ./bin/acxrun -np 8 --timeout 100 build/tests/ring_partitioned 2>&1 | tail -n 12
echo "### new tests"
for t in ordering stress; do timeout 120 ./bin/acxrun -np 2 --timeout 100 build/tests/$t 2>&1 | tail -n 3; done
echo "### allpairs graph np=2 (side streams)"
ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/p2p_bench allpairs --bytes 16777216 2>&1 | tail -n 1
echo "### bench.py N=2 (python)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 2>&1 | grep -E '^\{|rror' | tail -n 2 | cut -c1-1800
echo "### regcache test"
timeout 100 ./bin/acxrun -np 2 --timeout 90 build/tests/regcache 2>&1 | tail -n 2
timeout 100 ./bin/acxrun -np 2 --timeout 90 build/tests/large_msgs 16777216 2>&1 | tail -n 2
echo "### pingpong small/medium: knobs"
for cfg in "ACX_PDL=0 ACX_FUSE_RNDV=0" "ACX_PDL=1 ACX_FUSE_RNDV=0" "ACX_PDL=0 ACX_FUSE_RNDV=1" "ACX_PDL=1 ACX_FUSE_RNDV=1"; do
  echo "-- $cfg"; env $cfg ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/p2p_bench pingpong --min 8 --max 4194304 2>&1 | tail -n 1 | python -c "$summ"
done
echo "### pingpong private cudaMalloc buffers (registration cache)"
ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/p2p_bench pingpong --min 65536 --max 268435456 --no-heap 2>&1 | tail -n 1 | python -c "$summ"
echo "### pingpong large: movers"
for cfg in "ACX_MOVER=ldst" "ACX_MOVER=ldst256" "ACX_MOVER=tma"; do
  echo "-- $cfg"; env $cfg ACX_HEAP_BYTES=3g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/p2p_bench pingpong --min 16777216 --max 1073741824 2>&1 | tail -n 1 | python -c "$summ"
done
echo "### ring: movers"
for cfg in "ACX_MOVER=ldst" "ACX_MOVER=ldst256" "ACX_MOVER=tma"; do
  echo "-- $cfg"; env $cfg ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | cut -c1-230
done
echo "### halo api ours"
timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/halo_bench api --n 512 --steps 10 2>&1 | tail -n 1 | cut -c1-260
echo "### ppgemm api ours"
ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/ppgemm_bench api --steps 5 2>&1 | tail -n 3
echo "### trace: 8-byte pingpong timeline"
ACX_TRACE=1 ACX_TRACE_DIR=$O timeout 60 ./bin/acxrun -np 2 --timeout 50 bin/p2p_bench pingpong --min 8 --max 8 --iters 20 > /dev/null 2>&1; ls $O/acx_trace_rank*.jsonl
