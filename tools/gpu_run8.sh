#!/bin/bash
# 8-GPU pass: conformance at np=8, the four benchmark configs on both arms, bench.py N=8.
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
R=baseline/_ref/bin
NP=${NP:-8}
echo "### ctests np=$NP"
for t in ring_all_device ring_partitioned ring_all_graph; do timeout 90 ./bin/acxrun -np $NP --timeout 80 build/tests/$t 2>&1 | tail -n 1; done
timeout 90 ./bin/acxrun -np $NP --timeout 80 build/tests/large_msgs 16777216 2>&1 | tail -n 1
echo "### ring 256 MiB"
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | tee $O/ring_acx_n$NP.json
ACX_MOVER=tma ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | tee $O/ring_acx_tma_n$NP.json
timeout 90 ./bin/acxrun -np $NP --timeout 80 $R/p2p_bench ring --bytes 268435456 --no-heap 2>&1 | tail -n 1 | tee $O/ring_ref_n$NP.json
echo "### allpairs graph 16 MiB per pair"
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/p2p_bench allpairs --bytes 16777216 2>&1 | tail -n 1 | tee $O/allpairs_acx_n$NP.json
timeout 120 ./bin/acxrun -np $NP --timeout 110 $R/p2p_bench allpairs --bytes 16777216 --no-heap 2>&1 | tail -n 2 | tee $O/allpairs_ref_n$NP.json
echo "### halo 512^3 per rank"
timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1 | tee $O/halo_fused_n$NP.json
timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/halo_bench api --n 512 --steps 10 2>&1 | tail -n 1 | tee $O/halo_api_acx_n$NP.json
timeout 120 ./bin/acxrun -np $NP --timeout 110 $R/halo_bench api --n 512 --steps 10 --no-heap 2>&1 | tail -n 2 | tee $O/halo_api_ref_n$NP.json
echo "### ppgemm 8192^2 fp8, 4 partitions"
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/ppgemm_bench fused --steps 16 2>&1 | tail -n 2 | tee $O/ppgemm_fused_n$NP.json
timeout 150 ./bin/acxrun -np $NP --timeout 140 $R/ppgemm_bench api --steps 8 2>&1 | tail -n 2 | tee $O/ppgemm_api_ref_n$NP.json
echo "### bench.py N=$NP"
for impl in reference acx; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NP --impl $impl > $O/bench_${impl}_n$NP.json 2> $O/bench_${impl}_n$NP.err
  echo "bench $impl rc=$?"; grep '^{' $O/bench_${impl}_n$NP.json | tail -n 1 | cut -c1-1500; tail -n 2 $O/bench_${impl}_n$NP.err | cut -c1-300
done
