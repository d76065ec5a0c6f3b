#!/bin/bash
# Final 8-GPU pass: np=8 conformance incl. the new tests, re-measurement of the configs touched since the first pass.
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
R=baseline/_ref/bin
NP=8
echo "### ctests np=$NP"
for t in ring_partitioned ordering regcache; do timeout 90 ./bin/acxrun -np $NP --timeout 80 build/tests/$t 2>&1 | tail -n 2; done
timeout 90 ./bin/acxrun -np $NP --timeout 80 build/tests/stress 300 2>&1 | tail -n 2
echo "### ring 256 MiB"
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | tee $O/ring_acx_n${NP}_b.json
timeout 90 ./bin/acxrun -np $NP --timeout 80 $R/p2p_bench ring --bytes 268435456 --no-heap 2>&1 | tail -n 1 | tee $O/ring_ref_n${NP}_b.json
timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/p2p_bench ring --bytes 268435456 --no-heap 2>&1 | tail -n 1 | tee $O/ring_acx_private_n${NP}.json
echo "### allpairs graph 16 MiB per pair"
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/p2p_bench allpairs --bytes 16777216 2>&1 | tail -n 1 | tee $O/allpairs_acx_n${NP}_b.json
timeout 120 ./bin/acxrun -np $NP --timeout 110 $R/p2p_bench allpairs --bytes 16777216 --no-heap 2>&1 | tail -n 1 | tee $O/allpairs_ref_n${NP}_b.json
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/p2p_bench allpairs --bytes 1048576 2>&1 | tail -n 1
timeout 120 ./bin/acxrun -np $NP --timeout 110 $R/p2p_bench allpairs --bytes 1048576 --no-heap 2>&1 | tail -n 1
echo "### halo 512^3 per rank"
timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1 | tee $O/halo_fused_n${NP}_b.json
timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/halo_bench api --n 512 --steps 10 2>&1 | tail -n 1 | tee $O/halo_api_acx_n${NP}_b.json
echo "### ppgemm"
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/ppgemm_bench fused --steps 16 2>&1 | tail -n 1 | tee $O/ppgemm_fused_n${NP}_b.json
ACX_HEAP_BYTES=1g timeout 90 ./bin/acxrun -np $NP --timeout 80 bin/ppgemm_bench api --steps 8 2>&1 | tail -n 1 | tee $O/ppgemm_api_acx_n${NP}.json
echo "### bench.py"
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n > $O/bench_acx_n${n}_b.json 2> $O/bench_acx_n${n}_b.err
  echo "bench acx n=$n rc=$?"; grep '^{' $O/bench_acx_n${n}_b.json | tail -n 1 | cut -c1-1700; tail -n 2 $O/bench_acx_n${n}_b.err | cut -c1-300
done
