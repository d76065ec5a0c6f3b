"""acx-b200: GPU-initiated point-to-point communication for Blackwell (sm_100a).

The product is the C/CUDA runtime (include/mpi-acx.h, lib/libacx.so); this package is the
Python surface used by bench.py, the tests and torch users.
"""
from .runtime import (ANY_SOURCE, ANY_TAG, PROC_NULL, AcxError, PartitionedRequest, Request, barrier, empty, finalize, init, irecv, is_device_transport, isend,
                      precv_init, psend_init, rank, recv, send, sendrecv, size, startall, synchronize, stats, strided_layout, trace_dump, wait, wait_enqueue, waitall, waitall_enqueue)

__all__ = ["ANY_SOURCE", "ANY_TAG", "PROC_NULL", "AcxError", "PartitionedRequest", "Request", "barrier", "empty", "finalize", "init", "irecv", "is_device_transport", "isend",
           "precv_init", "psend_init", "rank", "recv", "send", "sendrecv", "size", "startall", "synchronize", "stats", "strided_layout", "trace_dump", "wait", "wait_enqueue", "waitall", "waitall_enqueue"]
__version__ = "0.1.0"
