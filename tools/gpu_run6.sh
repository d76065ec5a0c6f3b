#!/bin/bash
# 1-GPU pass: pytest, halo variants, ncu of the new halo kernel.
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
echo "### halo variants np=1 (512^3)"
for v in 2 3 4; do ACX_HALO_VARIANT=$v timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/halo_bench fused --n 512 --steps 20 2>&1 | tail -n 1 | cut -c1-260; done
echo "### halo verify variant 3/4 (256^3)"
for v in 3 4; do ACX_HALO_VARIANT=$v timeout 100 ./bin/acxrun -np 1 --timeout 90 bin/halo_bench both --n 256 --steps 5 2>&1 | tail -n 1; done
ACX_HALO_VARIANT=3 timeout 100 ./bin/acxrun -np 2 --timeout 90 bin/halo_bench both --n 256 --steps 5 2>&1 | tail -n 1
echo "### pytest gpu"
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 12
echo "### ncu halo (new kernel)"
ncu --set full --clock-control none --import-source on -k regex:halo_stencil_step_rows -s 2 -c 1 -o $O/prof_halo_rows -f bin/halo_bench fused --n 512 --steps 3 --warmup 2 --no-graph > $O/ncu_halo_rows.log 2>&1; echo "rc=$?"
echo "### sanitizer (memcheck on ring + partitioned, np=1)"
for t in ring ring_partitioned; do timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 build/tests/$t > $O/memcheck_$t.log 2>&1; echo "memcheck $t rc=$? $(grep -E 'ERROR SUMMARY' $O/memcheck_$t.log | tail -n 1)"; done
