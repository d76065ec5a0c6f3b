#!/bin/bash
# 1 GPU: CTA-pair GEMM (cta_group::2) numerics + timing, device trace of the copy-engine ring step, full GPU suite, smoke.
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call12; mkdir -p $O
echo "### 1. CTA-pair GEMM"
ok=1
for shape in "256 256 128" "256 512 1024" "1024 1024 4096" "8192 8192 8192"; do
  echo "[pair $shape] $(ACX_GEMM_PAIR=1 timeout 60 build/tests/gemm_fp8_test $shape 2>&1 | tail -n 4 | tr '\n' ' ')"
  ACX_GEMM_PAIR=1 timeout 60 build/tests/gemm_fp8_test $shape > /dev/null 2>&1 || ok=0
done
for i in 1 2; do
echo "[single 8192^3] $(ACX_GEMM_PAIR=0 timeout 60 build/tests/gemm_fp8_test 8192 8192 8192 20 2>&1 | grep bench)"
echo "[pair   8192^3] $(ACX_GEMM_PAIR=1 timeout 60 build/tests/gemm_fp8_test 8192 8192 8192 20 2>&1 | grep bench)"
done
echo "pair ok=$ok"
if [ $ok = 1 ]; then
  ACX_GEMM_PAIR=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_fp8_pair -s 2 -c 1 -f -o $O/gemm_pair build/tests/gemm_fp8_test 8192 8192 8192 2 > $O/ncu_gemm_pair.log 2>&1; tail -n 2 $O/ncu_gemm_pair.log
  echo "[ppgemm fused pair np=1] $(ACX_GEMM_PAIR=1 ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np 1 --timeout 100 bin/ppgemm_bench fused --steps 20 2>&1 | tail -n 1)"
fi
echo "### 2. ring 256 MiB N=1 with device trace (copy engine / SM mover)"
for v in "ACX_CE_MIN=4194304" "ACX_CE_MIN=0"; do
  mkdir -p $O/trace_$v
  echo "[$v] $(env $v ACX_TRACE=1 ACX_TRACE_DIR=$O/trace_$v ACX_HEAP_BYTES=1g timeout 60 ./bin/acxrun -np 1 --timeout 50 bin/p2p_bench ring --bytes 268435456 --steps 6 --warmup 3 2>&1 | tail -n 1 | cut -c1-200)"
done
echo "### 3. full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 | tee $O/pytest.txt
echo "### 4. smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
echo "### done"
