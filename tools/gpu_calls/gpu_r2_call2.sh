#!/bin/bash
# Round 2, call 2 (1 GPU): memcheck of the failing status_errors, the TMA halo kernel (numerics, timing, ncu), bench N=1.
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call2; mkdir -p $O
echo "### 1. status_errors under compute-sanitizer (np=1)"
ACX_RANK=0 ACX_SIZE=1 timeout 300 compute-sanitizer --tool memcheck --target-processes all --print-limit 8 ./bin/acxrun -np 1 --timeout 250 build/tests/status_errors > $O/status_errors_memcheck.txt 2>&1
grep -E "Invalid|at 0x|by thread|Address|in /|=========     Host Frame|PASS|FAIL|ERROR SUMMARY" $O/status_errors_memcheck.txt | head -40
echo "### 2. halo: numerics (pytest) + fused timing + verify against the API-level path"
timeout 600 python -m pytest tests -m gpu -q -k "halo or tags or truncation" 2>&1 | tail -n 15
for n in 512; do
  ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench both --n $n --steps 20 2>&1 | tail -n 4
  for c in 1 2; do echo "ctas/sm=$c: $(ACX_HALO_CTAS_PER_SM=$c ACX_HEAP_BYTES=2g timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/halo_bench fused --n $n --steps 30 2>&1 | tail -n 1)"; done
done
echo "### 3. ncu of the halo kernel"
ACX_HEAP_BYTES=2g ACX_RANK=0 ACX_SIZE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo_stencil_tma -s 3 -c 1 -f -o $O/halo_tma bin/halo_bench fused --n 512 --steps 4 --warmup 2 --no-graph > $O/ncu_halo.log 2>&1; tail -n 3 $O/ncu_halo.log
echo "### 4. bench.py N=1 both arms"
for impl in reference acx; do
  timeout 600 python bench.py --gpus 1 --impl $impl > $O/bench_${impl}_n1.json 2> $O/bench_${impl}_n1.err
  echo "bench.py impl=$impl rc=$?"; grep '^{' $O/bench_${impl}_n1.json | tail -n 1 | cut -c1-2500; tail -n 3 $O/bench_${impl}_n1.err
done
echo "### done"
