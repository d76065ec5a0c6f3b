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


def allpairs_exchange(send_chunks: Sequence, recv_chunks: Sequence, tag: int = 0, stream=None) -> None:
    """Chunk k goes to rank (me+1+k) % n and chunk k is received from rank (me-1-k) % n."""
    r, n = rt.rank(), rt.size()
    q: List[rt.Request] = []
    for k, (s, d) in enumerate(zip(send_chunks, recv_chunks)):
        q.append(rt.irecv(d, (r - 1 - k) % n, tag, stream))
        q.append(rt.isend(s, (r + 1 + k) % n, tag, stream))
    rt.waitall_enqueue(q, stream)
