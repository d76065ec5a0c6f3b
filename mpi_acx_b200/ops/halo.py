"""Fused halo-exchange + 7-point stencil (src/kernels/halo_stencil.cu) as a Python op."""
from __future__ import annotations

import ctypes

from .. import _native as N
from ..runtime import AcxError, _stream_arg


class HaloStencil:
    """Periodic px*py*pz decomposition, bf16 block nx*ny*nz per rank (nx 256 or 512, ny % 8 == 0, nz % 32 == 0).
    One fused kernel per step: TMA box loads feed a shared-memory plane ring, tiles acquire the neighbours' arrival flags,
    compute, and store their face cells straight into the neighbours' ghost cells over NVLink (the fields live, with
    their ghost shell, in the symmetric heap: ACX_HEAP_BYTES must hold both)."""

    def __init__(self, nx: int, ny: int, nz: int, decomp=(1, 1, 1), c0: float = 0.4, c1: float = 0.1):
        L = N.lib()
        L.acx_halo_create.argtypes = [ctypes.c_int] * 6 + [ctypes.c_float, ctypes.c_float, ctypes.POINTER(ctypes.c_void_p)]
        L.acx_halo_field.restype = ctypes.c_void_p
        L.acx_halo_field.argtypes = [ctypes.c_void_p]
        L.acx_halo_field_pitch.restype = None
        L.acx_halo_field_pitch.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
        L.acx_halo_step.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.acx_halo_exchange_initial.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.acx_halo_destroy.argtypes = [ctypes.c_void_p]
        L.acx_halo_steps_done.restype = ctypes.c_ulonglong
        L.acx_halo_steps_done.argtypes = [ctypes.c_void_p]
        self.shape = (nz, ny, nx)
        self.c0, self.c1 = c0, c1
        self.h = ctypes.c_void_p()
        if L.acx_halo_create(nx, ny, nz, *decomp, c0, c1, ctypes.byref(self.h)) != 0:
            raise AcxError("acx_halo_create failed")

    def field(self):
        """Current field as a torch bf16 tensor [nz, ny, nx]: a strided VIEW of the op's padded device array (rows and
        planes are pitched: the ghost shell lives between them), so .copy_() / .float() / indexing work in place."""
        import torch
        L = N.lib()
        ptr = L.acx_halo_field(self.h)
        rp, pp = ctypes.c_longlong(0), ctypes.c_longlong(0)
        L.acx_halo_field_pitch(self.h, ctypes.byref(rp), ctypes.byref(pp))
        nz, ny, nx = self.shape
        span = (nz - 1) * pp.value + (ny - 1) * rp.value + nx      # cells from the interior origin to the last interior cell
        holder = type("Buf", (), {})()
        holder.__cuda_array_interface__ = {"shape": (span * 2,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        flat = torch.as_tensor(holder, device="cuda").view(torch.bfloat16)
        return torch.as_strided(flat, (nz, ny, nx), (pp.value, rp.value, 1))

    def step(self, nsteps: int = 1, stream=None) -> None:
        cs = _stream_arg(stream)
        if N.lib().acx_halo_step(self.h, nsteps, cs) != 0:
            raise AcxError("acx_halo_step failed")

    @property
    def steps_done(self) -> int:
        return int(N.lib().acx_halo_steps_done(self.h))

    def close(self) -> None:
        if self.h:
            N.lib().acx_halo_destroy(self.h)
            self.h = ctypes.c_void_p()


def stencil_reference(u, c0: float = 0.4, c1: float = 0.1, steps: int = 1):
    """Plain PyTorch fp32 reference of the same op on a periodic global grid (bf16 storage per step)."""
    import torch
    for _ in range(steps):
        f = u.float()
        s = ((torch.roll(f, 1, 2) + torch.roll(f, -1, 2)) + (torch.roll(f, 1, 1) + torch.roll(f, -1, 1))) + (torch.roll(f, 1, 0) + torch.roll(f, -1, 0))
        u = (c0 * f + c1 * s).to(torch.bfloat16)
    return u
