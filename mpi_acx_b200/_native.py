"""ctypes binding of lib/libacx.so (the C/CUDA runtime is the product; this is glue).

The library is built in-tree by ``__graft_entry__.build()`` / ``make`` and loaded from
``<repo>/lib``. There is deliberately no pure-Python fallback: if the shared object is
missing the import fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
LIB_PATH = REPO / "lib" / "libacx.so"
# CI without a GPU can point the binding at the software-device build (build/emu/libacx_emu.so, see tests/emu): the device
# transport then runs on CPU tensors. Never set on a GPU box.
if os.environ.get("ACX_LIB"):
    LIB_PATH = Path(os.environ["ACX_LIB"]).resolve()

MPI_COMM_WORLD = 1
MPI_THREAD_MULTIPLE = 3
MPI_BYTE = (4 << 8) | 1
MPI_ANY_SOURCE = -2
MPI_ANY_TAG = -1
MPI_PROC_NULL = -3
QUEUE_STREAM = 0
QUEUE_GRAPH = 1


class MPIStatus(ctypes.Structure):
    _fields_ = [("MPI_SOURCE", ctypes.c_int), ("MPI_TAG", ctypes.c_int), ("MPI_ERROR", ctypes.c_int), ("_bytes", ctypes.c_longlong)]


class AcxStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_ulonglong) for n in (
        "sends_enqueued", "recvs_enqueued", "waits_enqueued", "bytes_sent", "bytes_received", "eager_msgs",
        "rendezvous_msgs", "bounce_copies", "kernels_launched", "graph_nodes_added", "slots_in_use",
        "slots_high_water", "heap_bytes_in_use", "heap_bytes_high_water", "copy_engine_msgs", "overflow_sends")]


def build_native(force: bool = False) -> None:
    """Build lib/libacx.so, bin/acxrun and the fused-kernel library with make (sm_100a)."""
    if force or not LIB_PATH.exists():
        import fcntl
        with open(REPO / ".build.lock", "w") as lk:   # several ranks may import at once: build exactly once
            fcntl.flock(lk, fcntl.LOCK_EX)
            try:
                if force or not LIB_PATH.exists():
                    subprocess.run(["make", "-C", str(REPO), "-j8", "lib", "bin/acxrun"], check=True, capture_output=True)
            finally:
                fcntl.flock(lk, fcntl.LOCK_UN)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            try:
                build_native()   # nvcc is part of the image on both the CPU and the GPU box
            except Exception as ex:  # noqa: BLE001
                raise ImportError(f"{LIB_PATH} is missing and `make` failed ({ex}); run __graft_entry__.build() first") from ex
        _lib = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
        vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        P = ctypes.POINTER
        _lib.MPI_Init_thread.argtypes = [vp, vp, ci, P(ci)]
        _lib.MPI_Comm_rank.argtypes = [ci, P(ci)]
        _lib.MPI_Comm_size.argtypes = [ci, P(ci)]
        _lib.MPI_Barrier.argtypes = [ci]
        _lib.MPI_Initialized.argtypes = [P(ci)]
        _lib.MPI_Type_contiguous.argtypes = [ci, ci, P(ci)]
        _lib.MPI_Type_create_hvector.argtypes = [ci, ci, ctypes.c_longlong, ci, P(ci)]
        _lib.MPI_Type_commit.argtypes = [P(ci)]
        _lib.MPI_Type_free.argtypes = [P(ci)]
        _lib.MPIX_Isend_enqueue.argtypes = [vp, ci, ci, ci, ci, ci, P(vp), ci, vp]
        _lib.MPIX_Irecv_enqueue.argtypes = [vp, ci, ci, ci, ci, ci, P(vp), ci, vp]
        _lib.MPIX_Wait_enqueue.argtypes = [P(vp), P(MPIStatus), ci, vp]
        _lib.MPIX_Send_enqueue.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp]
        _lib.MPIX_Recv_enqueue.argtypes = [vp, ci, ci, ci, ci, ci, P(MPIStatus), ci, vp]
        _lib.MPIX_Sendrecv_enqueue.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, ci, P(MPIStatus), ci, vp]
        _lib.MPIX_Waitall_enqueue.argtypes = [ci, P(vp), P(MPIStatus), ci, vp]
        _lib.MPIX_Wait.argtypes = [P(vp), P(MPIStatus)]
        _lib.MPIX_Waitall.argtypes = [ci, P(vp), P(MPIStatus)]
        _lib.MPIX_Psend_init.argtypes = [vp, ci, ctypes.c_longlong, ci, ci, ci, ci, ci, P(vp)]
        _lib.MPIX_Precv_init.argtypes = [vp, ci, ctypes.c_longlong, ci, ci, ci, ci, ci, P(vp)]
        _lib.MPIX_Prequest_create.argtypes = [vp, P(vp)]
        _lib.MPIX_Prequest_free.argtypes = [P(vp)]
        _lib.MPIX_Start.argtypes = [P(vp)]
        _lib.MPIX_Startall.argtypes = [ci, P(vp)]
        _lib.MPIX_Start_enqueue.argtypes = [P(vp), ci, vp]
        _lib.MPIX_Pwait_enqueue.argtypes = [P(vp), ci, vp]
        _lib.MPIX_Request_free.argtypes = [P(vp)]
        _lib.MPIX_Pready_host.argtypes = [ci, P(vp)]
        _lib.MPIX_Parrived_host.argtypes = [P(vp), ci, P(ci)]
        _lib.MPIX_Alloc_mem.argtypes = [cs, ci, P(vp)]
        _lib.MPIX_Free_mem.argtypes = [vp]
        _lib.MPIX_Get_stats.argtypes = [P(AcxStats)]
        _lib.MPIX_Trace_dump.argtypes = [ctypes.c_char_p]
    return _lib
