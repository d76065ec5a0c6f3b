"""A chain of fp8 linear stages, one per rank (pipeline parallelism), BASELINE config #4."""
from __future__ import annotations

from ..ops.gemm import PipelineGemm, gemm_fp8_reference
from .. import runtime as rt


class PipelineLinearChain:
    def __init__(self, batch: int = 8192, width: int = 8192, partitions: int = 4, scale: float | None = None):
        self.scale = scale if scale is not None else 1.5 / (width ** 0.5)
        self.op = PipelineGemm(batch, width, partitions, self.scale)
        self.stage = rt.rank()

    def init_weights(self, seed: int = 0):
        import torch
        g = torch.Generator(device="cuda").manual_seed(1000 + seed + self.stage)
        w = (torch.randn(self.op.D, self.op.D, device="cuda", generator=g)).clamp_(-3, 3).to(torch.float8_e4m3fn)
        self.op.weights().copy_(w)
        return w

    def forward_microbatch(self, stream=None):
        self.op.step(stream)

    def reference_stage(self, x):
        return gemm_fp8_reference(x, self.op.weights(), self.scale)

    def close(self):
        self.op.close()
