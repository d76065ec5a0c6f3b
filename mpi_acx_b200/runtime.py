"""Process-level runtime: init/finalize, requests, symmetric tensors, stream-ordered p2p.

Python mirror of the C API in include/mpi-acx.h (itself signature-compatible with the
reference's include/mpi-acx.h:42-104). One process per GPU; ranks come from torchrun's
RANK/WORLD_SIZE/LOCAL_RANK (or acxrun's ACX_* variables).
"""
from __future__ import annotations

import atexit
import ctypes
import os
from typing import Iterable, Optional, Sequence

from . import _native as N

_state = {"init": False, "rank": 0, "size": 1}


class AcxError(RuntimeError):
    pass


def _ck(rc: int, what: str) -> None:
    if rc != 0:
        raise AcxError(f"{what} failed with code {rc}")


def init(device: Optional[int] = None) -> None:
    """MPI_Init_thread + cudaSetDevice (LOCAL_RANK) + MPIX_Init. Idempotent."""
    if _state["init"]:
        return
    L = N.lib()
    import torch
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("ACX_LOCAL_RANK", "0")))
    if torch.cuda.is_available() and not os.environ.get("ACX_FORCE_HOST"):
        torch.cuda.set_device(device % torch.cuda.device_count())
        torch.cuda.current_stream()  # create the primary context before MPIX_Init allocates
        torch.zeros(1, device="cuda")
    prov = ctypes.c_int(0)
    _ck(L.MPI_Init_thread(None, None, N.MPI_THREAD_MULTIPLE, ctypes.byref(prov)), "MPI_Init_thread")
    _ck(L.MPIX_Init(), "MPIX_Init")
    r, s = ctypes.c_int(0), ctypes.c_int(0)
    L.MPI_Comm_rank(N.MPI_COMM_WORLD, ctypes.byref(r))
    L.MPI_Comm_size(N.MPI_COMM_WORLD, ctypes.byref(s))
    _state.update(init=True, rank=r.value, size=s.value)
    atexit.register(finalize)


def finalize() -> None:
    if not _state["init"]:
        return
    L = N.lib()
    _state["init"] = False
    L.MPIX_Finalize()
    L.MPI_Finalize()


def rank() -> int:
    return _state["rank"]


def size() -> int:
    return _state["size"]


def is_device_transport() -> bool:
    return bool(N.lib().MPIX_Is_device_transport())


def synchronize() -> None:
    """Wait for all enqueued work of this rank (torch.cuda.synchronize(); on the software device: its cudaDeviceSynchronize)."""
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    elif is_device_transport():
        N.lib().cudaDeviceSynchronize()


def barrier() -> None:
    N.lib().MPI_Barrier(N.MPI_COMM_WORLD)


def stats() -> dict:
    s = N.AcxStats()
    N.lib().MPIX_Get_stats(ctypes.byref(s))
    return {k: int(getattr(s, k)) for k, _ in s._fields_}


# ---- symmetric heap tensors ---------------------------------------------------------------
class _CudaBuffer:
    """Minimal __cuda_array_interface__ carrier so torch can wrap symmetric-heap memory."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}

    def __del__(self):
        try:
            if _state["init"] and self.ptr:
                N.lib().MPIX_Free_mem(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


def empty(shape: Sequence[int] | int, dtype=None):
    """Tensor backed by the symmetric communication heap (zero-copy send/recv)."""
    import torch
    dtype = dtype or torch.uint8
    if isinstance(shape, int):
        shape = (shape,)
    numel = 1
    for d in shape:
        numel *= int(d)
    nbytes = max(numel * torch.empty((), dtype=dtype).element_size(), 16)
    if not is_device_transport():
        return torch.empty(tuple(shape), dtype=dtype)
    p = ctypes.c_void_p()
    _ck(N.lib().MPIX_Alloc_mem(nbytes, 0, ctypes.byref(p)), "MPIX_Alloc_mem")
    buf = _CudaBuffer(p.value, nbytes)
    if not torch.cuda.is_available():   # software device (ACX_LIB=...libacx_emu.so): "device" memory is host-addressable
        t = torch.frombuffer((ctypes.c_ubyte * nbytes).from_address(p.value), dtype=torch.uint8)
    else:
        t = torch.as_tensor(buf, device="cuda")
    t = t[: numel * torch.empty((), dtype=dtype).element_size()].view(dtype).view(tuple(shape))
    t._acx_keepalive = buf  # tie heap block lifetime to the tensor
    return t


# ---- requests -----------------------------------------------------------------------------
class Request:
    __slots__ = ("h", "keep")

    def __init__(self, handle: ctypes.c_void_p, keep=None):
        self.h = handle
        self.keep = keep


def _stream_arg(stream):
    import torch
    if stream is None:
        stream = torch.cuda.current_stream() if torch.cuda.is_available() and is_device_transport() else None
    cs = ctypes.c_void_p(stream.cuda_stream if stream is not None else 0)
    return cs


def _buf(t):
    if not t.is_contiguous():
        raise AcxError("tensors must be contiguous")
    return ctypes.c_void_p(t.data_ptr()), t.numel() * t.element_size()


# Strided tensor views (a face of a 3-D block, a column of a matrix, ...) travel without a .contiguous() copy in user
# code: the view's sizes/strides become an MPI derived datatype (nested hvectors around one contiguous run) and the
# library's pack / unpack kernels gather and scatter it. The wire format is the view's elements in logical (row-major)
# order, so a strided send matches a contiguous receive of the same shape and vice versa.
_type_cache: dict = {}


def strided_layout(shape: Sequence[int], strides: Sequence[int], itemsize: int):
    """(run_bytes, [(count, stride_bytes), ...] inner -> outer) of a tensor view; raises if it cannot be expressed."""
    dims = [(int(n), int(st) * itemsize) for n, st in zip(shape, strides) if int(n) != 1]
    if any(st < 0 for _, st in dims):
        raise AcxError("negative strides are not supported")
    if any(st == 0 for _, st in dims):
        raise AcxError("expanded (stride 0) views are not supported")
    run = itemsize
    while dims and dims[-1][1] == run:      # fold the innermost dimensions that continue the run
        run *= dims[-1][0]
        dims.pop()
    folded: list = []
    for n, st in reversed(dims):            # inner -> outer; merge adjacent dimensions that form one regular stride
        if folded and st == folded[-1][0] * folded[-1][1]:
            folded[-1] = (folded[-1][0] * n, folded[-1][1])
        else:
            folded.append((n, st))
    if len(folded) > 3:
        raise AcxError("view needs more than 3 strided dimensions; call .contiguous() first")
    return run, folded


def _typed_buf(t):
    """(pointer, count, MPI datatype) for any dense or regularly strided tensor view."""
    if t.is_contiguous():
        p, nb = _buf(t)
        return p, nb, N.MPI_BYTE
    run, dims = strided_layout(t.shape, t.stride(), t.element_size())
    key = (run, tuple(dims))
    dt = _type_cache.get(key)
    if dt is None:
        L = N.lib()
        cur = ctypes.c_int(0)
        _ck(L.MPI_Type_contiguous(run, N.MPI_BYTE, ctypes.byref(cur)), "MPI_Type_contiguous")
        for n, st in dims:
            nxt = ctypes.c_int(0)
            _ck(L.MPI_Type_create_hvector(n, 1, st, cur.value, ctypes.byref(nxt)), "MPI_Type_create_hvector")
            cur = nxt
        _ck(L.MPI_Type_commit(ctypes.byref(cur)), "MPI_Type_commit")
        dt = _type_cache[key] = cur.value
    return ctypes.c_void_p(t.data_ptr()), 1, dt


PROC_NULL, ANY_SOURCE, ANY_TAG = N.MPI_PROC_NULL, N.MPI_ANY_SOURCE, N.MPI_ANY_TAG


def isend(t, dest: int, tag: int = 0, stream=None) -> Request:
    """Stream-ordered send of a tensor (dense, or a regularly strided view). `dest` may be PROC_NULL."""
    p, n, dt = _typed_buf(t)
    h, cs = ctypes.c_void_p(), _stream_arg(stream)
    _ck(N.lib().MPIX_Isend_enqueue(p, n, dt, dest, tag, N.MPI_COMM_WORLD, ctypes.byref(h), N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Isend_enqueue")
    return Request(h, t)


def irecv(t, source: int, tag: int = 0, stream=None) -> Request:
    """Stream-ordered receive into a tensor (dense, or a regularly strided view). `source` may be PROC_NULL
    (or ANY_SOURCE / tag ANY_TAG where the transport supports wildcards)."""
    p, n, dt = _typed_buf(t)
    h, cs = ctypes.c_void_p(), _stream_arg(stream)
    _ck(N.lib().MPIX_Irecv_enqueue(p, n, dt, source, tag, N.MPI_COMM_WORLD, ctypes.byref(h), N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Irecv_enqueue")
    return Request(h, t)


def send(t, dest: int, tag: int = 0, stream=None) -> None:
    """Blocking-on-the-stream send (MPIX_Send_enqueue): later work on the stream runs once the tensor may be reused.
    An eager-size message is a single trigger kernel."""
    p, n, dt = _typed_buf(t)
    cs = _stream_arg(stream)
    _ck(N.lib().MPIX_Send_enqueue(p, n, dt, dest, tag, N.MPI_COMM_WORLD, N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Send_enqueue")


def recv(t, source: int, tag: int = 0, stream=None) -> None:
    """Blocking-on-the-stream receive (MPIX_Recv_enqueue): later work on the stream sees the data. An eager-size dense
    receive is ONE kernel (post + wait fused) instead of the two of irecv + wait_enqueue."""
    p, n, dt = _typed_buf(t)
    cs = _stream_arg(stream)
    _ck(N.lib().MPIX_Recv_enqueue(p, n, dt, source, tag, N.MPI_COMM_WORLD, None, N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Recv_enqueue")


def sendrecv(send_t, dest: int, recv_t, source: int, tag: int = 0, stream=None) -> None:
    """MPI_Sendrecv on the stream (MPIX_Sendrecv_enqueue): one call per exchange step; with small dense tensors on both
    sides, one kernel (receive posted, send deposited, receive completed)."""
    sp, sn, sdt = _typed_buf(send_t)
    rp, rn, rdt = _typed_buf(recv_t)
    cs = _stream_arg(stream)
    _ck(N.lib().MPIX_Sendrecv_enqueue(sp, sn, sdt, dest, tag, rp, rn, rdt, source, tag, N.MPI_COMM_WORLD, None, N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Sendrecv_enqueue")


def waitall_enqueue(reqs: Iterable[Request], stream=None) -> None:
    reqs = list(reqs)
    arr = (ctypes.c_void_p * len(reqs))(*[r.h for r in reqs])
    cs = _stream_arg(stream)
    _ck(N.lib().MPIX_Waitall_enqueue(len(reqs), arr, None, N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Waitall_enqueue")
    for r in reqs:
        r.h = ctypes.c_void_p()


def wait_enqueue(req: Request, stream=None) -> None:
    waitall_enqueue([req], stream)


def wait(req: Request) -> dict:
    st = N.MPIStatus()
    _ck(N.lib().MPIX_Wait(ctypes.byref(req.h), ctypes.byref(st)), "MPIX_Wait")
    return {"source": st.MPI_SOURCE, "tag": st.MPI_TAG, "error": st.MPI_ERROR, "bytes": st._bytes}


def waitall(reqs: Iterable[Request]) -> list:
    """Host-blocking MPIX_Waitall (include/mpi-acx.h; the reference's loop over MPIX_Wait, src/sendrecv.cu:642-651):
    returns one status dict per request and nulls the handles."""
    reqs = list(reqs)
    arr = (ctypes.c_void_p * len(reqs))(*[r.h for r in reqs])
    sts = (N.MPIStatus * len(reqs))()
    _ck(N.lib().MPIX_Waitall(len(reqs), arr, sts), "MPIX_Waitall")
    for r in reqs:
        r.h = ctypes.c_void_p()
    return [{"source": st.MPI_SOURCE, "tag": st.MPI_TAG, "error": st.MPI_ERROR, "bytes": st._bytes} for st in sts]


def trace_dump(path: str) -> None:
    """Write this rank's device trace ring (ACX_TRACE=1) as JSON lines; tools/trace_to_chrome.py merges the ranks."""
    _ck(N.lib().MPIX_Trace_dump(str(path).encode()), "MPIX_Trace_dump")


# ---- partitioned ----------------------------------------------------------------------------
class PartitionedRequest:
    """Persistent partitioned channel endpoint; `.prequest` is the device handle (int) that
    CUDA kernels pass to acx::pready / acx::parrived (include/acx_device.cuh)."""

    def __init__(self, t, partitions: int, peer: int, tag: int, send: bool):
        p, nb = _buf(t)
        if nb % partitions:
            raise AcxError("tensor bytes must divide evenly into partitions")
        self.h = ctypes.c_void_p()
        self.keep = t
        self.partitions = partitions
        fn = N.lib().MPIX_Psend_init if send else N.lib().MPIX_Precv_init
        _ck(fn(p, partitions, nb // partitions, N.MPI_BYTE, peer, tag, N.MPI_COMM_WORLD, 0, ctypes.byref(self.h)), "MPIX_P*_init")
        self._preq = None

    @property
    def prequest(self) -> int:
        if self._preq is None:
            q = ctypes.c_void_p()
            _ck(N.lib().MPIX_Prequest_create(self.h, ctypes.byref(q)), "MPIX_Prequest_create")
            self._preq = q.value
        return self._preq

    def start(self) -> None:
        _ck(N.lib().MPIX_Start(ctypes.byref(self.h)), "MPIX_Start")

    def start_enqueue(self, stream=None) -> None:
        cs = _stream_arg(stream)
        _ck(N.lib().MPIX_Start_enqueue(ctypes.byref(self.h), N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Start_enqueue")

    def wait(self) -> None:
        _ck(N.lib().MPIX_Wait(ctypes.byref(self.h), None), "MPIX_Wait")

    def wait_enqueue(self, stream=None) -> None:
        cs = _stream_arg(stream)
        _ck(N.lib().MPIX_Pwait_enqueue(ctypes.byref(self.h), N.QUEUE_STREAM, ctypes.byref(cs)), "MPIX_Pwait_enqueue")

    def pready(self, partition: int) -> None:
        _ck(N.lib().MPIX_Pready_host(partition, ctypes.byref(self.h)), "MPIX_Pready")

    def parrived(self, partition: int) -> bool:
        f = ctypes.c_int(0)
        _ck(N.lib().MPIX_Parrived_host(ctypes.byref(self.h), partition, ctypes.byref(f)), "MPIX_Parrived")
        return bool(f.value)

    def free(self) -> None:
        if self.h:
            _ck(N.lib().MPIX_Request_free(ctypes.byref(self.h)), "MPIX_Request_free")


def startall(reqs: Sequence[PartitionedRequest]) -> None:
    arr = (ctypes.c_void_p * len(reqs))(*[r.h for r in reqs])
    _ck(N.lib().MPIX_Startall(len(reqs), arr), "MPIX_Startall")


def psend_init(t, partitions: int, dest: int, tag: int = 0) -> PartitionedRequest:
    return PartitionedRequest(t, partitions, dest, tag, True)


def precv_init(t, partitions: int, source: int, tag: int = 0) -> PartitionedRequest:
    return PartitionedRequest(t, partitions, source, tag, False)
