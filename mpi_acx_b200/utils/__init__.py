from .clocks import ClockSampler
from .timing import max_over_ranks, time_cuda

__all__ = ["ClockSampler", "time_cuda", "max_over_ranks"]
