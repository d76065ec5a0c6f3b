from .gemm import PipelineGemm, gemm_fp8, gemm_fp8_reference
from .halo import HaloStencil, stencil_reference

__all__ = ["HaloStencil", "stencil_reference", "gemm_fp8", "gemm_fp8_reference", "PipelineGemm"]
