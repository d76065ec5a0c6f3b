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
print("PY-EMU OK", r, acx.stats()["kernels_launched"])
acx.finalize()
