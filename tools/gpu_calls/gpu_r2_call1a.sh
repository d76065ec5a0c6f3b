#!/bin/bash
# Round 2, 1-GPU call: full GPU suite without -x, self-copy probe, bench.py N=1 both arms.
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call1a; mkdir -p $O
top() { awk '{for(i=1;i<=NF;i++) if($i=="GB/s") print $(i-1), $0}' $1 | sort -rn | head -${2:-10} | cut -d' ' -f2-; }
lscpu | grep -i "numa\|^CPU(s)\|Model name"; nvidia-smi topo -m > $O/topo.txt 2>&1
echo "### A. pytest -m gpu (no -x)"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 40 | tee $O/pytest.txt
echo "### B. bwprobe self"
timeout 200 tools/bwprobe 1 > $O/bwprobe_self1.txt 2>&1; grep -c GB/s $O/bwprobe_self1.txt; grep "copy engine" $O/bwprobe_self1.txt; top $O/bwprobe_self1.txt 10
echo "### E. bench.py N=1 both arms"
for impl in reference acx; do
  timeout 400 python bench.py --gpus 1 --impl $impl > $O/bench_${impl}_n1.json 2> $O/bench_${impl}_n1.err
  echo "bench.py impl=$impl rc=$?"; grep '^{' $O/bench_${impl}_n1.json | tail -n 1 | cut -c1-1800; tail -n 3 $O/bench_${impl}_n1.err
done
ACX_BENCH_E2E_CHUNKS=1 timeout 400 python bench.py --gpus 1 --no-extras 2>&1 | grep '^{' | cut -c1-900
echo "### done"
