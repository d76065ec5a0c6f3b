"""Ring exchange of torch tensors through the Python binding.
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 examples/torch_ring.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mpi_acx_b200 as acx  # noqa: E402
from mpi_acx_b200.parallel import ring_exchange  # noqa: E402

acx.init()
r, n = acx.rank(), acx.size()
dev = "cuda" if acx.is_device_transport() else "cpu"
send = acx.empty((1024, 1024), torch.bfloat16)        # symmetric heap: zero-copy over NVLink
recv = acx.empty((1024, 1024), torch.bfloat16)
send.fill_(r + 1)
ring_exchange(send, recv, tag=7)                       # enqueued on the current stream; returns immediately
if dev == "cuda":
    torch.cuda.synchronize()
print(f"rank {r}: received tensor filled with {recv[0, 0].item()} from rank {(r - 1) % n}")
acx.finalize()
