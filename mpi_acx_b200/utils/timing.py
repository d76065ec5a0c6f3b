"""CUDA-event timing helpers (device time on the launching stream, max over ranks)."""
from __future__ import annotations


def time_cuda(fn, iters: int, warmup: int = 3, stream=None) -> float:
    """Milliseconds per call of fn(), measured with CUDA events after `warmup` untimed calls."""
    import torch
    stream = stream or torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
