"""Host-staged exchange: inputs that start in pinned host memory are sent while they are still arriving.

The reference hands a host buffer to MPI whole: the H2D copy (PCIe, ~55 GB/s) and the transfer (NVLink) run one after
the other. Here the message is cut into chunks; chunk i is copied to the device on a copy stream while chunk i-1 is
already on the wire (one stream-triggered send per chunk, gated on that chunk's copy by an event). The step then costs
the H2D time plus ONE chunk's transfer instead of H2D + the whole transfer. Everything is the public stream API
(isend / irecv / waitall_enqueue); this module only arranges the calls.
"""
from __future__ import annotations

from typing import List, Optional

from .. import runtime as rt


def pinned_empty(nbytes: int):
    """Pinned host staging buffer (uint8). Call after acx.init(): MPIX_Init binds the process to the GPU's NUMA node, so
    the pages land on the socket the GPU hangs off (see acx_numa_bind in src/core/acx.cu)."""
    import torch
    t = torch.empty(int(nbytes), dtype=torch.uint8)
    return t.pin_memory() if torch.cuda.is_available() and rt.is_device_transport() else t


class StagedExchange:
    """Persistent plan for `send host_src -> dest` / `recv <- source` with `chunks`-way H2D/transfer overlap.

    send_dev / recv_dev are device tensors (symmetric-heap tensors from acx.empty() make the transfer zero-copy) of the
    same byte size; tags tag_base .. tag_base+chunks-1 are used."""

    def __init__(self, send_dev, recv_dev, dest: int, source: int, chunks: int = 8, tag_base: int = 32):
        import torch
        self.send_dev, self.recv_dev = send_dev.view(torch.uint8).reshape(-1), recv_dev.view(torch.uint8).reshape(-1)
        n = self.send_dev.numel()
        assert self.recv_dev.numel() == n
        chunks = max(1, min(int(chunks), 16, n // 4096 or 1))
        per = -(-n // chunks)
        per = (per + 4095) // 4096 * 4096
        self.bounds = [(lo, min(n, lo + per)) for lo in range(0, n, per)]
        self.dest, self.source, self.tag_base = dest, source, tag_base
        self.cuda = torch.cuda.is_available() and rt.is_device_transport()
        if self.cuda:
            self.copy_stream = torch.cuda.Stream()
            self.ev_copied = [torch.cuda.Event() for _ in self.bounds]
            self.ev_step = torch.cuda.Event()
        self._first = True

    def step(self, host_src, stream=None) -> None:
        """Enqueue one exchange; stream work enqueued afterwards sees recv_dev filled. host_src: pinned uint8 tensor."""
        import torch
        if not self.cuda:   # host transport / software device: no copy engine to overlap with
            self.send_dev.copy_(host_src)
            q = [rt.irecv(self.recv_dev, self.source, self.tag_base, stream), rt.isend(self.send_dev, self.dest, self.tag_base, stream)]
            rt.waitall_enqueue(q, stream)
            return
        main = stream or torch.cuda.current_stream()
        # the staging buffer may be overwritten only after the previous step's sends have completed (its waitall ran on `main`)
        self.ev_step.record(main)
        self.copy_stream.wait_event(self.ev_step)
        q: List[rt.Request] = [rt.irecv(self.recv_dev[lo:hi], self.source, self.tag_base + i, main) for i, (lo, hi) in enumerate(self.bounds)]
        for i, (lo, hi) in enumerate(self.bounds):
            with torch.cuda.stream(self.copy_stream):
                self.send_dev[lo:hi].copy_(host_src[lo:hi], non_blocking=True)
                self.ev_copied[i].record(self.copy_stream)
            main.wait_event(self.ev_copied[i])
            q.append(rt.isend(self.send_dev[lo:hi], self.dest, self.tag_base + i, main))
        rt.waitall_enqueue(q, main)


def staged_ring_exchange(plan: Optional[StagedExchange], host_src, stream=None) -> None:
    plan.step(host_src, stream)
