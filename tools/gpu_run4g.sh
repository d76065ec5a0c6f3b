#!/bin/bash
export ACX_WATCHDOG_MS=30000
for cfg in "ACX_SIDE_STREAMS=1 ACX_MOVER_MAX_CTAS=148" "ACX_SIDE_STREAMS=0" "ACX_SIDE_STREAMS=1 ACX_MOVER_MAX_CTAS=74"; do
  echo "-- python bench N=4: $cfg"
  env $cfg timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --no-extras 2>/dev/null | grep '^{' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'])"
done
for cfg in "ACX_SIDE_STREAMS=1 ACX_MOVER_MAX_CTAS=148" "ACX_SIDE_STREAMS=0"; do
  echo "-- C ring np=4: $cfg"; env $cfg ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 4 --timeout 50 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | cut -c1-200
  echo "-- C allpairs np=4: $cfg"; env $cfg ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 4 --timeout 50 bin/p2p_bench allpairs --bytes 16777216 2>&1 | tail -n 1 | cut -c1-200
done
