"""Neighbour-exchange patterns on the stream-triggered API (the pattern all six reference tests use,
e.g. test/src/ring.c:78-81), plus an all-pairs exchange (BASELINE config #5)."""
from __future__ import annotations

from typing import List, Sequence

from .. import runtime as rt


def ring_neighbours():
    r, n = rt.rank(), rt.size()
    return (r - 1) % n, (r + 1) % n


def ring_exchange(send, recv, tag: int = 0, stream=None) -> None:
    """send -> right neighbour, recv <- left neighbour; returns once the ops are enqueued on `stream`
    (stream work enqueued afterwards sees the received data)."""
    left, right = ring_neighbours()
    q = [rt.irecv(recv, left, tag, stream), rt.isend(send, right, tag, stream)]
    rt.waitall_enqueue(q, stream)


def ring_exchange_prepost(send, recv_bufs, steps: int, tag: int = 0, stream=None, before_send=None, after_recv=None) -> None:
    """`steps` ring exchanges with the receive of step i+1 posted BEFORE step i is waited on (two receive
    buffers, alternating). Pre-posting is what MPI applications do for bandwidth: every send then finds its
    receive already posted, so the sender moves the data itself (one NVLink push per GPU per step, no
    arrival-order lottery about which GPU does the copying). Step i lands in recv_bufs[i % 2].
    `before_send(i)` / `after_recv(i, buf)` are enqueued around each step (e.g. H2D staging, D2H read-back)."""
    left, right = ring_neighbours()
    pending = rt.irecv(recv_bufs[0], left, tag, stream)
    for i in range(steps):
        if before_send is not None:
            before_send(i)
        rs = rt.isend(send, right, tag, stream)
        nxt = rt.irecv(recv_bufs[(i + 1) % 2], left, tag, stream) if i + 1 < steps else None
        rt.waitall_enqueue([pending, rs], stream)
        if after_recv is not None:
            after_recv(i, recv_bufs[i % 2])
        pending = nxt


def allpairs_exchange(send_chunks: Sequence, recv_chunks: Sequence, tag: int = 0, stream=None) -> None:
    """Chunk k goes to rank (me+1+k) % n and chunk k is received from rank (me-1-k) % n."""
    r, n = rt.rank(), rt.size()
    q: List[rt.Request] = []
    for k, (s, d) in enumerate(zip(send_chunks, recv_chunks)):
        q.append(rt.irecv(d, (r - 1 - k) % n, tag, stream))
        q.append(rt.isend(s, (r + 1 + k) % n, tag, stream))
    rt.waitall_enqueue(q, stream)
