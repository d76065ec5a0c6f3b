#!/bin/bash
export ACX_WATCHDOG_MS=20000
R=baseline/_ref/bin
for t in ring_all_device ring_all_graph ordering regcache; do timeout 40 ./bin/acxrun -np 2 --timeout 30 build/tests/$t 2>&1 | tail -n 1; done
timeout 40 ./bin/acxrun -np 2 --timeout 30 build/tests/large_msgs 16777216 2>&1 | tail -n 1
timeout 40 ./bin/acxrun -np 2 --timeout 30 build/tests/stress 300 2>&1 | tail -n 1
ACX_HEAP_BYTES=1g timeout 40 ./bin/acxrun -np 2 --timeout 30 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | cut -c1-200
ACX_HEAP_BYTES=1g timeout 40 ./bin/acxrun -np 2 --timeout 30 bin/p2p_bench allpairs --bytes 16777216 2>&1 | tail -n 1 | cut -c1-200
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --no-extras 2>/dev/null | grep '^{' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('pybench', j['value'], j['ms_per_step'], j['e2e']['value'], j['verified'])"
ACX_HEAP_BYTES=1g timeout 40 ./bin/acxrun -np 2 --timeout 30 bin/p2p_bench pingpong_graph --min 8 --max 2097152 --iters 20 2>&1 | tail -n 1 | cut -c1-700
timeout 60 ./bin/acxrun -np 2 --timeout 50 $R/p2p_bench pingpong_graph --min 8 --max 2097152 --iters 20 --no-heap 2>&1 | tail -n 1 | cut -c1-700
