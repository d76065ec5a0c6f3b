"""Stage-to-stage activation hand-off with partitioned requests (host-driven rounds; the fused,
host-free variant is ops.PipelineGemm)."""
from __future__ import annotations

from .. import runtime as rt


class PartitionedLink:
    """Persistent partitioned channel between this stage and the next / previous one."""

    def __init__(self, send_buf=None, recv_buf=None, partitions: int = 4, tag: int = 0):
        r, n = rt.rank(), rt.size()
        self.send = rt.psend_init(send_buf, partitions, r + 1, tag) if send_buf is not None and r + 1 < n else None
        self.recv = rt.precv_init(recv_buf, partitions, r - 1, tag) if recv_buf is not None and r > 0 else None

    def start(self) -> None:
        active = [x for x in (self.recv, self.send) if x is not None]
        if active:
            rt.startall(active)

    def wait(self) -> None:
        for x in (self.send, self.recv):
            if x is not None:
                x.wait()

    def free(self) -> None:
        for x in (self.send, self.recv):
            if x is not None:
                x.free()
