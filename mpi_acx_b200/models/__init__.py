from .pipeline_mlp import PipelineLinearChain
from .stencil import Diffusion3D, default_decomp

__all__ = ["Diffusion3D", "default_decomp", "PipelineLinearChain"]
