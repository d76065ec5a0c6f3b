#!/bin/bash
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
R=baseline/_ref/bin
echo "### ppgemm fused"
ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/ppgemm_bench fused --steps 10 2>&1 | tail -n 3
ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/ppgemm_bench fused --steps 10 2>&1 | tail -n 3 | tee $O/ppgemm_fused_n2.json
echo "### halo variants np=1"
for v in 0 1 2; do ACX_HALO_VARIANT=$v timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/halo_bench fused --n 512 --steps 20 2>&1 | tail -n 1; done
echo "### halo verify + np=2"
timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/halo_bench both --n 256 --steps 5 2>&1 | tail -n 3
for v in 1 2; do ACX_HALO_VARIANT=$v timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/halo_bench fused --n 512 --steps 20 2>&1 | tail -n 1; done
echo "### ring movers"
for cfg in "ACX_MOVER=ldst" "ACX_MOVER=ldst ACX_MOVER_MAX_CTAS=148" "ACX_MOVER=ldst ACX_MOVER_MAX_CTAS=296" "ACX_MOVER=tma" "ACX_MOVER=tma ACX_MOVER_MAX_CTAS=74" "ACX_MOVER=tma ACX_MOVER_MAX_CTAS=32"; do
  echo "-- $cfg"; env $cfg ACX_HEAP_BYTES=1g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/p2p_bench ring --bytes 268435456 2>&1 | tail -n 1 | cut -c1-230
done
echo "### pingpong large, movers"
for cfg in "ACX_MOVER=ldst" "ACX_MOVER=tma"; do
  echo "-- $cfg"; env $cfg ACX_HEAP_BYTES=3g timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/p2p_bench pingpong --min 1048576 --max 1073741824 2>&1 | tail -n 1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print([(r['bytes']>>20, round(r['oneway_us'],1), round(r['gbs'],1)) for r in j['rows']])"
done
echo "### large_msgs with tma mover"
ACX_MOVER=tma timeout 100 ./bin/acxrun -np 2 --timeout 90 build/tests/large_msgs 2>&1 | tail -n 2
echo "### reference api benches (small, bounded)"
timeout 100 ./bin/acxrun -np 2 --timeout 90 $R/halo_bench api --n 256 --steps 5 --no-heap 2>&1 | tail -n 2
timeout 100 ./bin/acxrun -np 2 --timeout 90 $R/halo_bench api --n 512 --steps 10 --no-heap 2>&1 | tail -n 2 | tee $O/halo_api_ref_n2.json
timeout 100 ./bin/acxrun -np 2 --timeout 90 $R/ppgemm_bench api --steps 5 2>&1 | tail -n 2 | tee $O/ppgemm_api_ref_n2.json
timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/ppgemm_bench api --steps 5 2>&1 | tail -n 2 | tee $O/ppgemm_api_acx_n2.json
echo "### pytest gpu"
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 12
