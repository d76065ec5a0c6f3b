#!/bin/bash
export ACX_WATCHDOG_MS=30000
O=gpurun_out; mkdir -p $O
R=baseline/_ref/bin
echo "### ctests np=2 (2 GPUs)"
tools/gpu_ctests.sh "2" > /dev/null 2>&1; grep -E "^==|FAIL|WATCHDOG" $O/ctests.log
echo "### pingpong acx heap"
ACX_HEAP_BYTES=3g timeout 300 ./bin/acxrun -np 2 --timeout 280 bin/p2p_bench pingpong --max 1073741824 > $O/pingpong_acx_heap.json 2> $O/pingpong_acx_heap.err; echo rc=$?
echo "### pingpong acx private"
ACX_HEAP_BYTES=3g timeout 300 ./bin/acxrun -np 2 --timeout 280 bin/p2p_bench pingpong --max 268435456 --no-heap > $O/pingpong_acx_private.json 2> $O/pingpong_acx_private.err; echo rc=$?; tail -n 2 $O/pingpong_acx_private.err
echo "### pingpong reference"
timeout 400 ./bin/acxrun -np 2 --timeout 380 $R/p2p_bench pingpong --max 1073741824 --iters 40 > $O/pingpong_ref.json 2> $O/pingpong_ref.err; echo rc=$?; tail -n 3 $O/pingpong_ref.err
echo "### ring 256 MiB"
ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench ring --bytes 268435456 | tee $O/ring_acx.json
timeout 120 ./bin/acxrun -np 2 --timeout 100 $R/p2p_bench ring --bytes 268435456 --no-heap | tee $O/ring_ref.json
echo "### allpairs graph 16 MiB"
ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/p2p_bench allpairs --bytes 16777216 | tee $O/allpairs_acx.json
timeout 120 ./bin/acxrun -np 2 --timeout 100 $R/p2p_bench allpairs --bytes 16777216 --no-heap 2>&1 | tail -n 3 | tee $O/allpairs_ref.json
echo "### halo verify (256^3, np=1 and np=2)"
timeout 120 ./bin/acxrun -np 1 --timeout 100 bin/halo_bench both --n 256 --steps 5 2>&1 | tail -n 4
timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/halo_bench both --n 256 --steps 5 2>&1 | tail -n 4
echo "### halo 512^3 per rank"
timeout 120 ./bin/acxrun -np 1 --timeout 100 bin/halo_bench fused --n 512 --steps 50 2>&1 | tail -n 2 | tee $O/halo_fused_n1.json
timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/halo_bench fused --n 512 --steps 50 2>&1 | tail -n 2 | tee $O/halo_fused_n2.json
timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/halo_bench fused --n 512 --steps 50 --no-graph 2>&1 | tail -n 2
timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/halo_bench api --n 512 --steps 20 2>&1 | tail -n 2 | tee $O/halo_api_acx_n2.json
timeout 200 ./bin/acxrun -np 2 --timeout 180 $R/halo_bench api --n 512 --steps 20 --no-heap 2>&1 | tail -n 3 | tee $O/halo_api_ref_n2.json
echo "### bench.py reference n=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --impl reference 2>&1 | grep -E '^\{|rror' | tail -n 3
python - <<'PY'
import json
for f in ["pingpong_acx_heap.json","pingpong_acx_private.json","pingpong_ref.json"]:
    try:
        s=open("gpurun_out/"+f).read()
        if not s.rstrip().endswith("]}"): s=s.rstrip().rstrip(",")+"]}"
        j=json.loads(s)
        print(f, [(r["bytes"], round(r["oneway_us"],1), round(r["gbs"],1)) for r in j["rows"] if r["bytes"] in (8,1024,8192,16384,65536,1<<20,16<<20,256<<20,1<<30)])
    except Exception as e: print(f, "ERR", e)
PY
echo "### gemm fp8 single GPU"
timeout 120 build/tests/gemm_fp8_test 1024 1024 1024 5 2>&1 | tail -n 12
timeout 120 build/tests/gemm_fp8_test 8192 8192 8192 10 2>&1 | tail -n 6
echo "### ppgemm fused np=2"
ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np 2 --timeout 100 bin/ppgemm_bench fused --steps 10 2>&1 | tail -n 6 | tee $O/ppgemm_fused_n2.json
ACX_HEAP_BYTES=1g timeout 120 ./bin/acxrun -np 1 --timeout 100 bin/ppgemm_bench fused --steps 10 2>&1 | tail -n 4
timeout 300 ./bin/acxrun -np 2 --timeout 280 $R/ppgemm_bench api --steps 5 2>&1 | tail -n 4 | tee $O/ppgemm_api_ref_n2.json
