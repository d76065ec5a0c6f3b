"""Python binding on the software CUDA device (ACX_LIB=build/emu/libacx_emu.so): strided tensor views through derived
datatypes (pack / unpack kernels on the DEVICE transport), symmetric-heap tensors, blocking send / recv — on CPU tensors.
Run by tests/test_host.py::test_python_binding_on_software_device under acxrun."""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mpi_acx_b200 as acx
acx.init(0)
assert acx.is_device_transport()
r, n = acx.rank(), acx.size()
left, right = (r - 1) % n, (r + 1) % n
for shape, dtype in (((8, 8, 8), torch.float32), ((64, 48, 40), torch.bfloat16), ((24, 16, 32), torch.float32)):
    torch.manual_seed(7)
    base = (torch.randn(shape) * 10).to(dtype)
    blk, ghost = base.clone(), torch.zeros(shape, dtype=dtype)
    views = [(blk[:, :, -2], ghost[:, :, 0]), (blk[:, -2, :], ghost[:, 0, :]), (blk[-2], ghost[0]), (blk[1:5, ::2, 3:9], ghost[1:5, ::2, 3:9])]
    for tag, (sv, rv) in enumerate(views):
        dense = torch.zeros(sv.shape, dtype=dtype)
        q = [acx.irecv(rv, left, 20 + tag), acx.irecv(dense, left, 40 + tag), acx.isend(sv, right, 20 + tag), acx.isend(sv, right, 40 + tag)]
        acx.waitall_enqueue(q)
        acx.synchronize()
        assert torch.equal(rv, sv) and torch.equal(dense, sv.contiguous()), (shape, tag)   # every rank holds the same seeded data
    touched = torch.zeros(shape, dtype=dtype)
    touched[:, :, 0] = base[:, :, -2]; touched[:, 0, :] = base[:, -2, :]; touched[0] = base[-2]; touched[1:5, ::2, 3:9] = base[1:5, ::2, 3:9]
    assert torch.equal(ghost, touched), shape
# symmetric-heap tensors, large (rendezvous) and small, blocking calls
for numel in (5, 3000, 200000):
    s, d = acx.empty(numel, torch.float32), acx.empty(numel, torch.float32)
    s.copy_(torch.arange(numel, dtype=torch.float32) + r); d.zero_()
    if n == 1 or r % 2 == 0:
        q = acx.irecv(d, left, 3); acx.send(s, right, 3); acx.wait_enqueue(q)
    else:
        acx.recv(d, left, 3); acx.send(s, right, 3)
    acx.synchronize()
    assert torch.equal(d, torch.arange(numel, dtype=torch.float32) + left), numel
# one-call exchange step (one fused kernel for small dense tensors)
a_, b_ = torch.full((100,), float(r)), torch.zeros(100)
k0 = acx.stats()["kernels_launched"]
acx.sendrecv(a_, right, b_, left, tag=8)
acx.synchronize()
assert bool((b_ == float(left)).all()) and acx.stats()["kernels_launched"] - k0 == 1
# one-call exchange step (one fused kernel for small dense tensors)
a_, b_ = torch.full((100,), float(r)), torch.zeros(100)
k0 = acx.stats()["kernels_launched"]
acx.sendrecv(a_, right, b_, left, tag=8)
acx.synchronize()
assert bool((b_ == float(left)).all()) and acx.stats()["kernels_launched"] - k0 == 1
# parallel/: pre-posted double-buffered ring (what bench.py runs from 3 GPUs up), all-pairs exchange, partitioned pipeline link
from mpi_acx_b200.parallel import ring as pring
from mpi_acx_b200.parallel.pipeline import PartitionedLink
sbuf, rbufs = acx.empty(50000, torch.int32), [acx.empty(50000, torch.int32), acx.empty(50000, torch.int32)]
seen = []
def before(i): sbuf.fill_(1000 * r + i)
def after(i, buf): acx.synchronize(); seen.append(int(buf[0]) == 1000 * left + i and int(buf[-1]) == 1000 * left + i)
pring.ring_exchange_prepost(sbuf, rbufs, steps=6, tag=5, before_send=lambda i: (acx.synchronize(), before(i)), after_recv=after)
acx.synchronize()
assert seen == [True] * 6, seen
chunks_s = [torch.full((3000,), 100 * r + k, dtype=torch.int32) for k in range(n)]
chunks_r = [torch.zeros(3000, dtype=torch.int32) for _ in range(n)]
pring.allpairs_exchange(chunks_s, chunks_r, tag=6)
acx.synchronize()
for k in range(n):
    assert int(chunks_r[k][0]) == 100 * ((r - 1 - k) % n) + k, (k, int(chunks_r[k][0]))
if n > 1:   # activations flow rank 0 -> 1 -> ... over persistent partitioned channels, host-driven rounds
    act_in, act_out = acx.empty(4 * 256, torch.float32), acx.empty(4 * 256, torch.float32)
    link = PartitionedLink(send_buf=act_out, recv_buf=act_in, partitions=4, tag=7)
    for rnd in range(3):
        act_out.fill_(float(10 * r + rnd)); act_in.zero_()
        link.start()
        if link.send is not None:
            for p_ in range(4): link.send.pready(p_)
        link.wait()
        if r > 0:
            assert bool((act_in == float(10 * (r - 1) + rnd)).all()), rnd
    link.free()
# host-blocking waitall with statuses: eager and rendezvous sizes, heap and private buffers
for nel, mk in ((16, torch.zeros), (40000, torch.zeros), (40000, lambda k, dtype: acx.empty(k, dtype))):
    sb, rb = mk(nel, dtype=torch.int32), mk(nel, dtype=torch.int32)
    sb.fill_(7 + r); rb.zero_()
    sts = acx.waitall([acx.irecv(rb, left, tag=8), acx.isend(sb, (r + 1) % n, tag=8)])
    assert sts[0]["source"] == left and sts[0]["tag"] == 8 and sts[0]["bytes"] == 4 * nel and sts[0]["error"] == 0, sts
    assert bool((rb == 7 + left).all())
# a burst of sends before the receives they need (more than a ring holds): the triggers beyond the ring wait on pool streams and
# one waitall completes everything
ov0 = acx.stats()["overflow_sends"]
outs = [torch.full((6 + k % 5,), 100 * r + k, dtype=torch.int64) for k in range(45)]
ins = [torch.zeros(6 + k % 5, dtype=torch.int64) for k in range(45)]
q = [acx.isend(t, right, tag=12) for t in outs] + [acx.irecv(t, left, tag=12) for t in ins]
acx.waitall_enqueue(q)
acx.synchronize()
assert all(bool((ins[k] == 100 * left + k).all()) for k in range(45))
assert acx.stats()["overflow_sends"] - ov0 == 45 - 32, acx.stats()["overflow_sends"] - ov0
print("PY-EMU OK", r, acx.stats()["kernels_launched"])
acx.finalize()
