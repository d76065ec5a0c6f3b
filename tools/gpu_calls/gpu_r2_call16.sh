#!/bin/bash
# 1 GPU, last call of the round: TMA-store epilogue of the GEMM, 16 x 16 halo tiles, the full GPU suite, smoke, bench.py N=1.
export ACX_WATCHDOG_MS=30000
O=gpurun_out/call16; mkdir -p $O
echo "### 1. GEMM with the TMA-store epilogue"
for pair in 1 0; do for shape in "256 512 1024" "1024 1024 4096" "8192 8192 8192"; do
  echo "[pair=$pair tma_store $shape] $(ACX_GEMM_PAIR=$pair ACX_GEMM_TMA_STORE=1 timeout 60 build/tests/gemm_fp8_test $shape 10 2>&1 | tail -n 3 | tr '\n' ' ' | cut -c1-330)"
done; done
echo "[pair=1 plain 8192] $(ACX_GEMM_PAIR=1 timeout 60 build/tests/gemm_fp8_test 8192 8192 8192 20 2>&1 | grep bench)"
echo "[pair=1 tma   8192] $(ACX_GEMM_PAIR=1 ACX_GEMM_TMA_STORE=1 timeout 60 build/tests/gemm_fp8_test 8192 8192 8192 20 2>&1 | grep bench)"
echo "[ppgemm np=2 on one GPU, TMA store into the peer mapping] $(ACX_GEMM_TMA_STORE=2 ACX_HEAP_BYTES=1g timeout 150 ./bin/acxrun -np 2 --timeout 120 bin/ppgemm_bench fused --m 2048 --d 2048 --steps 6 2>&1 | tail -n 1 | cut -c1-300)"
echo "[ppgemm np=2 on one GPU, plain] $(ACX_HEAP_BYTES=1g timeout 150 ./bin/acxrun -np 2 --timeout 120 bin/ppgemm_bench fused --m 2048 --d 2048 --steps 6 2>&1 | tail -n 1 | cut -c1-300)"
echo "### 2. halo 512^3: 8 x 32 tiles vs 16 x 16 tiles"
hb() { env "$@" ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench fused --n 512 --steps 30 2>&1 | tail -n 1 | grep -o '"ms_per_step": [0-9.e-]*\|"max_abs_err[a-z_]*": [0-9.e-]*' | tr '\n' ' '; }
for rep in 1 2; do for v in "ACX_HALO_TILE=8" "ACX_HALO_TILE=16"; do echo "[$v] $(hb $v)"; done; done | tee $O/halo.txt
echo "[both 16] $(ACX_HALO_TILE=16 ACX_HEAP_BYTES=2g timeout 200 ./bin/acxrun -np 1 --timeout 180 bin/halo_bench both --n 512 --steps 10 2>&1 | tail -n 1 | cut -c1-400)"
echo "### 3. full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 | tee $O/pytest.txt
echo "### 4. smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1 | cut -c1-200
echo "### 5. bench.py N=1 (acx arm, with the secondary numbers)"
timeout 900 python bench.py > $O/bench_acx_n1.json 2> $O/bench_acx_n1.err; echo "rc=$?"; grep '^{' $O/bench_acx_n1.json | tail -n 1 | cut -c1-3000
echo "### done"
