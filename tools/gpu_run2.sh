#!/bin/bash
# 2-GPU measurement pass: reference sanity, ping-pong sweep, ring, bench.py both arms, pytest.
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
R=baseline/_ref/bin
echo "### reference's own tests on the mini-MPI shim"
for t in ring ring-all ring-all-device ring-partitioned; do timeout 60 ./bin/acxrun -np 2 --timeout 50 $R/$t > $O/ref_$t.log 2>&1; echo "ref $t rc=$?"; tail -n 2 $O/ref_$t.log; done
echo "### pingpong acx heap"
ACX_HEAP_BYTES=3g timeout 300 ./bin/acxrun -np 2 --timeout 280 bin/p2p_bench pingpong --max 1073741824 > $O/pingpong_acx_heap.json 2> $O/pingpong_acx_heap.err; echo rc=$?; cut -c1-1500 $O/pingpong_acx_heap.json
echo "### pingpong acx private cudaMalloc buffers (bounced)"
ACX_HEAP_BYTES=3g timeout 300 ./bin/acxrun -np 2 --timeout 280 bin/p2p_bench pingpong --max 268435456 --no-heap > $O/pingpong_acx_private.json 2> $O/pingpong_acx_private.err; echo rc=$?; cut -c1-1500 $O/pingpong_acx_private.json
echo "### pingpong reference"
timeout 400 ./bin/acxrun -np 2 --timeout 380 $R/p2p_bench pingpong --max 1073741824 --iters 40 > $O/pingpong_ref.json 2> $O/pingpong_ref.err; echo rc=$?; cut -c1-1500 $O/pingpong_ref.json; tail -n 3 $O/pingpong_ref.err
echo "### ring 256 MiB"
ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench ring --bytes 268435456 | tee $O/ring_acx.json
timeout 120 ./bin/acxrun -np 2 --timeout 100 $R/p2p_bench ring --bytes 268435456 --no-heap | tee $O/ring_ref.json
echo "### allpairs graph 16 MiB"
ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench allpairs --bytes 16777216 | tee $O/allpairs_acx.json
timeout 120 ./bin/acxrun -np 2 --timeout 100 $R/p2p_bench allpairs --bytes 16777216 --no-heap | tee $O/allpairs_ref.json
echo "### bench.py"
for n in 1 2; do
  for impl in acx reference; do
    if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --impl $impl > $O/bench_${impl}_n$n.json 2> $O/bench_${impl}_n$n.err
    else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --impl $impl > $O/bench_${impl}_n$n.json 2> $O/bench_${impl}_n$n.err; fi
    echo "bench $impl n=$n rc=$?"; grep '^{' $O/bench_${impl}_n$n.json | tail -n 1; tail -n 3 $O/bench_${impl}_n$n.err
  done
done
echo "### pytest gpu"
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 15 | tee $O/pytest_gpu.log
