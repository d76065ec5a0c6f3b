"""tcgen05/TMEM fp8 GEMM (src/kernels/gemm_fp8_sm100.cuh) and the pipeline-parallel chain built on it."""
from __future__ import annotations

import ctypes

from .. import _native as N
from ..runtime import AcxError, _stream_arg


def gemm_fp8(x, w, scale: float = 1.0, out=None, stream=None):
    """y[M,N] (float8_e4m3fn) = scale * x[M,K] @ w[N,K]^T, fp32 accumulation in TMEM."""
    import torch
    assert x.dtype == torch.float8_e4m3fn and w.dtype == torch.float8_e4m3fn and x.is_contiguous() and w.is_contiguous()
    M, K = x.shape
    Nn, K2 = w.shape
    assert K == K2 and M % 128 == 0 and Nn % 256 == 0 and K % 128 == 0, "tile alignment: M%128, N%256, K%128"
    if out is None:
        out = torch.empty((M, Nn), device=x.device, dtype=torch.float8_e4m3fn)
    L = N.lib()
    L.acx_gemm_fp8.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    rc = L.acx_gemm_fp8(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, Nn, K, scale, _stream_arg(stream))
    if rc:
        raise AcxError(f"acx_gemm_fp8 failed ({rc})")
    return out


def gemm_fp8_reference(x, w, scale: float = 1.0):
    """Plain PyTorch fp32 reference of the same op."""
    import torch
    return ((x.float() @ w.float().t()) * scale).to(torch.float8_e4m3fn)


class PipelineGemm:
    """One stage per rank; activations flow stage->stage in `partitions` row panels, written by the producing
    GEMM's epilogue into the next GPU's HBM and consumed by the next GEMM's TMA producer as they arrive."""

    def __init__(self, M: int, D: int, partitions: int = 4, scale: float = 1.0):
        L = N.lib()
        L.acx_ppgemm_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.POINTER(ctypes.c_void_p)]
        for f in ("acx_ppgemm_weights", "acx_ppgemm_input", "acx_ppgemm_output"):
            getattr(L, f).restype = ctypes.c_void_p
        L.acx_ppgemm_weights.argtypes = [ctypes.c_void_p]
        L.acx_ppgemm_input.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.acx_ppgemm_output.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.acx_ppgemm_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.acx_ppgemm_destroy.argtypes = [ctypes.c_void_p]
        L.acx_ppgemm_microbatches.restype = ctypes.c_ulonglong
        L.acx_ppgemm_microbatches.argtypes = [ctypes.c_void_p]
        self.M, self.D = M, D
        self.h = ctypes.c_void_p()
        if L.acx_ppgemm_create(M, D, partitions, scale, ctypes.byref(self.h)) != 0:
            raise AcxError("acx_ppgemm_create failed")

    def _wrap(self, ptr, shape):
        import torch
        nbytes = shape[0] * shape[1]
        holder = type("Buf", (), {})()
        holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(holder, device="cuda").view(torch.float8_e4m3fn).view(shape)

    def weights(self):
        return self._wrap(N.lib().acx_ppgemm_weights(self.h), (self.D, self.D))

    def input(self, buf: int = 0):
        return self._wrap(N.lib().acx_ppgemm_input(self.h, buf), (self.M, self.D))

    def output(self, buf: int = 0):
        return self._wrap(N.lib().acx_ppgemm_output(self.h, buf), (self.M, self.D))

    def step(self, stream=None) -> None:
        if N.lib().acx_ppgemm_step(self.h, _stream_arg(stream)) != 0:
            raise AcxError("acx_ppgemm_step failed")

    @property
    def microbatches(self) -> int:
        return int(N.lib().acx_ppgemm_microbatches(self.h))

    def close(self) -> None:
        if self.h:
            N.lib().acx_ppgemm_destroy(self.h)
            self.h = ctypes.c_void_p()
