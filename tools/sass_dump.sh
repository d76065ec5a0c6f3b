#!/bin/bash
# Per-kernel SASS listings + resource usage of the hot kernels in lib/libacx.so -> profiles/sass/ (no GPU needed).
cd "$(dirname "$0")/.."
out=profiles/sass; mkdir -p $out
cuobjdump -sass lib/libacx.so > /tmp/acx_all.sass
cuobjdump --dump-resource-usage lib/libacx.so > /tmp/acx_res.txt 2>/dev/null
for k in mover_plain mover_tma send_trigger recv_trigger wait_kernel sendrecv_fused halo_stencil_tmaE halo_stencil_tma16 gemm_fp8_kernel gemm_fp8_pair_kernel ce_gate ce_complete; do
  f=$k; [ $k = halo_stencil_tmaE ] && f=halo_stencil_tma   # the mangled name's length suffix tells the 8 x 32 kernel from halo_stencil_tma16
  awk -v k="$k" '/Function : /{f = index($0, k) > 0} f{print}' /tmp/acx_all.sass > $out/$f.sass
  n=$(grep -c "^\s*/\*[0-9a-f]\{4\}\*/" $out/$f.sass)
  res=$(grep -A1 "Function.*$k" /tmp/acx_res.txt | grep REG | head -1 | sed 's/^ *//')
  echo "$k: $n instructions; $res"
done | tee $out/INDEX.txt
python3 tools/sass_census.py lib/libacx.so > $out/census.txt 2>/dev/null || true
