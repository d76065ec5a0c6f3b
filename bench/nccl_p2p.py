#!/usr/bin/env python3
"""NCCL point-to-point baseline (BASELINE.md configs #2 and #5 name `ncclSend/ncclRecv` as the independent baseline
next to the reference-proxy build): the same three traffic patterns as bench/p2p_bench.cu, driven through
torch.distributed's NCCL backend (`batch_isend_irecv` = one ncclGroupStart/End per step), device-timed with CUDA
events, max over ranks. Library code on purpose — this is a yardstick, not part of the product.

    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 bench/nccl_p2p.py ring --bytes 268435456
    torchrun ... bench/nccl_p2p.py pingpong                     # 8 B - 1 GiB sweep (2 ranks)
    torchrun ... bench/nccl_p2p.py allpairs --bytes 16777216 [--graph]

NCCL_P2P_DRYRUN_CPU=1 runs the control flow on gloo/CPU (wall-clock timing, not a measurement) for the CPU test suite.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

DRY = os.environ.get("NCCL_P2P_DRYRUN_CPU") == "1"


class Timer:
    """CUDA events on the current stream (device time); wall clock in the CPU dry run."""

    def __enter__(self):
        if DRY:
            self.t0 = time.perf_counter()
        else:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if DRY:
            self.ms = (time.perf_counter() - self.t0) * 1e3
        else:
            self.e1.record()
            torch.cuda.synchronize()
            self.ms = self.e0.elapsed_time(self.e1)


def max_over_ranks(x, dev):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def exchange(pairs):
    """pairs: list of ('send'|'recv', tensor, peer). One NCCL group."""
    ops = [dist.P2POp(dist.isend if k == "send" else dist.irecv, t, p) for k, t, p in pairs]
    for w in dist.batch_isend_irecv(ops):
        w.wait()          # stream-ordered on NCCL (no host block); blocks on gloo


def ring(a, rank, world, dev):
    n = a.bytes
    sbuf = torch.full((n,), (rank + 1) & 0xFF, dtype=torch.uint8, device=dev)
    rbuf = torch.zeros(n, dtype=torch.uint8, device=dev)
    right, left = (rank + 1) % world, (rank - 1) % world

    def step():
        if world == 1:
            rbuf.copy_(sbuf)
        else:
            exchange([("recv", rbuf, left), ("send", sbuf, right)])
    for _ in range(a.warmup):
        step()
    dist.barrier()
    with Timer() as t:
        for _ in range(a.steps):
            step()
    ms = max_over_ranks(t.ms, dev) / a.steps
    bad = int((rbuf != ((left + 1) & 0xFF)).sum().item())
    return {"bench": "ring", "impl": "nccl-p2p", "n_gpus": world, "bytes": n, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 6),
            "gbs_per_gpu": round(n / ms / 1e6, 3), "gbs_total": round(n * world / ms / 1e6, 3), "bad_words": bad}


def pingpong(a, rank, world, dev):
    assert world >= 2, "pingpong needs 2 ranks"
    rows = []
    size = 8
    while size <= a.max_bytes:
        buf = torch.zeros(size, dtype=torch.uint8, device=dev)
        iters = 200 if size <= (1 << 20) else 20
        def rt():
            if rank == 0:
                dist.send(buf, 1); dist.recv(buf, 1)
            elif rank == 1:
                dist.recv(buf, 0); dist.send(buf, 0)
        for _ in range(5):
            rt()
        dist.barrier()
        with Timer() as t:
            for _ in range(iters):
                rt()
        ms = max_over_ranks(t.ms if rank < 2 else 0.0, dev)
        oneway_us = ms * 1e3 / iters / 2
        rows.append({"bytes": size, "iters": iters, "oneway_us": round(oneway_us, 3), "gbs": round(size / oneway_us / 1e3, 3)})
        size *= 2
    return {"bench": "pingpong", "impl": "nccl-p2p", "rows": rows}


def allpairs(a, rank, world, dev):
    n = a.bytes
    sb = [torch.full((n,), rank + 1, dtype=torch.uint8, device=dev) for _ in range(world)]
    rb = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(world)]
    peers = [p for p in range(world) if p != rank]

    def step():
        if peers:
            exchange([("recv", rb[p], p) for p in peers] + [("send", sb[p], p) for p in peers])
    for _ in range(a.warmup):
        step()
    graph = None
    if a.graph and not DRY:
        # Same shape as the MPIX all-pairs benchmark: the step is captured once and replayed.
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                step()
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(a.warmup):
            graph.replay()
    dist.barrier()
    with Timer() as t:
        for _ in range(a.steps):
            graph.replay() if graph else step()
    ms = max_over_ranks(t.ms, dev) / a.steps
    bad = sum(int((rb[p] != p + 1).sum().item()) for p in peers)
    out = n * len(peers)
    return {"bench": "allpairs_graph" if graph else "allpairs", "impl": "nccl-p2p", "n_gpus": world, "bytes_per_pair": n, "steps": a.steps,
            "ms_per_step": round(ms, 6), "gbs_per_gpu_out": round(out / ms / 1e6, 3), "gbs_total": round(out * world / ms / 1e6, 3), "bad_words": bad}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["ring", "pingpong", "allpairs"])
    ap.add_argument("--bytes", type=int, default=256 << 20)
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if DRY:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {"ring": ring, "pingpong": pingpong, "allpairs": allpairs}[a.mode](a, rank, world, dev)
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
