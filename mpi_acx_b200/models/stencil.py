"""Jacobi-style 3-D diffusion on a periodic grid, domain-decomposed over all ranks; the flagship
"model" of the halo-exchange benchmark (BASELINE config #3)."""
from __future__ import annotations

from ..ops.halo import HaloStencil, stencil_reference
from .. import runtime as rt


def default_decomp(n: int):
    px = py = pz = 1
    d = 0
    while n > 1 and n % 2 == 0:
        if d % 3 == 0:
            px *= 2
        elif d % 3 == 1:
            py *= 2
        else:
            pz *= 2
        n //= 2
        d += 1
    return px * n, py, pz


class Diffusion3D:
    def __init__(self, block=(512, 512, 512), c0: float = 0.4, c1: float = 0.1, decomp=None):
        self.decomp = decomp or default_decomp(rt.size())
        nx, ny, nz = block
        self.op = HaloStencil(nx, ny, nz, self.decomp, c0, c1)
        self.c0, self.c1 = c0, c1

    @property
    def field(self):
        return self.op.field()

    def forward(self, steps: int = 1, stream=None):
        self.op.step(steps, stream)
        return self.op.field()

    __call__ = forward

    def reference(self, u0_global, steps: int):
        return stencil_reference(u0_global, self.c0, self.c1, steps)

    def close(self):
        self.op.close()
