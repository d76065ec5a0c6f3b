#!/usr/bin/env python
"""Headline benchmark (driver contract): ring send/recv bandwidth through the stream-triggered
MPIX_Isend_enqueue / MPIX_Irecv_enqueue / MPIX_Waitall_enqueue API on N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # acx-b200 (this repo)
    python bench.py --impl reference --gpus N ...            # unmodified NVIDIA/mpi-acx from baseline/_ref

Metric (BASELINE.json): "send/recv GB/s" — every rank sends MSG bytes to its right neighbour and
receives MSG bytes from its left one per step (N=1: to itself); value = total bytes delivered per
second over all N GPUs, device-timed with CUDA events, max over ranks. MSG = 256 MiB by default so
the 512 MiB working set per GPU is larger than the 126 MB L2 (no flush needed between steps).
`e2e` repeats the measurement with a host->device copy of the step's send buffer from pinned memory
and a device->host read-back of 4 KiB of the received data inside every timed step.
Extra keys carry the latency / small-message / partitioned numbers gathered in the same run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="acx", choices=["acx", "reference"])
    ap.add_argument("--bytes", type=int, default=256 << 20)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--order", default="simple", choices=["auto", "prepost", "simple"],
                    help="prepost: the receive of step i+1 is posted before step i is waited on (double-buffered); simple: irecv, isend, waitall per step; "
                         "auto: prepost from 3 GPUs up (where the arrival-order lottery of the simple order costs up to 2x, BASELINE.md 4.6), simple below")
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))



# ------------------------------------------------------------------------------------------
# Secondary numbers, same keys from BOTH arms: the shared C harnesses (bench/*.cu — one source, compiled once against
# libacx and once with -DACX_BENCH_REFERENCE=1 against the unmodified reference) run as one process per rank under
# the launcher's environment. Small-message ring steps, CUDA-graph all-pairs, API-level halo step, API-level
# pipeline GEMM; the acx arm adds its fused kernels.
def c_harness_extras(bindir: Path, world: int, local: int, rank: int, ref: bool, tag: str) -> dict:
    import time
    out: dict = {}
    env = dict(os.environ)
    env.setdefault("ACX_WATCHDOG_MS", "60000")
    env["MPIACX_NFLAGS"] = "8192"
    if not ref:
        env["ACX_HEAP_BYTES"] = str(1 << 30)
    noheap = ["--no-heap"] if ref else []
    port = os.environ.get("MASTER_PORT", "0")
    from mpi_acx_b200.utils.numa import gpu_numa_cpus
    cpus = gpu_numa_cpus(local) if os.environ.get("ACX_NUMA_BIND", "1") != "0" else None
    runs = []
    for nb in (8, 4096, 1 << 20):
        runs.append((f"ring_step_us_{nb}B", "p2p_bench", ["ring", "--bytes", str(nb), "--steps", "200", "--warmup", "20"] + noheap + ([] if ref else ["--sendrecv"]), lambda j: round(j["ms_per_step"] * 1e3, 3)))
    if world >= 2:
        for nb, name in ((1 << 20, "1MiB"), (16 << 20, "16MiB")):
            runs.append((f"allpairs_graph_{name}_gbs_per_gpu", "p2p_bench", ["allpairs", "--bytes", str(nb)] + noheap, lambda j: round(j["gbs_per_gpu_out"], 2)))
    runs.append(("halo_api_512cube_ms_per_step", "halo_bench", ["api", "--n", "512", "--steps", "10"], lambda j: round(j["ms_per_step"], 5)))
    runs.append(("ppgemm_api_fp8_8192_ms_per_microbatch", "ppgemm_bench", ["api", "--steps", "10"], lambda j: round(j["ms_per_step"], 5)))
    if not ref:
        runs.append(("halo_fused_512cube_ms_per_step", "halo_bench", ["fused", "--n", "512", "--steps", "30"], lambda j: round(j["ms_per_step"], 5)))
        runs.append(("ppgemm_fused_fp8_8192_ms_per_microbatch", "ppgemm_bench", ["fused", "--steps", "40"], lambda j: round(j["ms_per_step"], 5)))   # 40 microbatches: the pipeline's fill (stages - 1 hand-offs) is part of the timed region
    for i, (key, exe, args, pick) in enumerate(runs):
        env["ACX_JOBID"] = f"{tag}_{port}_{i}_{os.environ.get('TORCHELASTIC_RUN_ID', 'x')}"
        try:
            p = subprocess.run([str(bindir / exe)] + args, env=env, capture_output=True, text=True, timeout=300,
                               preexec_fn=(lambda: os.sched_setaffinity(0, cpus)) if cpus else None)
            if rank == 0:
                lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
                out[key] = pick(json.loads(lines[-1])) if p.returncode == 0 and lines else f"failed rc={p.returncode}: {(p.stderr or p.stdout)[-160:]!r}"
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                out[key] = f"failed: {ex!r}"[:200]
        time.sleep(0.2)
    return out

# ------------------------------------------------------------------------------------------
def run_reference(a) -> int:
    rank, world, local = env_rank()
    exe = REPO / "baseline" / "_ref" / "bin" / "p2p_bench"
    if not exe.exists() and Path("/root/reference/Makefile").exists() and rank == 0:
        subprocess.run([str(REPO / "baseline" / "build_ref.sh")], capture_output=True, timeout=900)   # one-off build of the unmodified reference
    if world > 1 and Path("/root/reference/Makefile").exists():
        import time
        for _ in range(600):   # other ranks wait for rank 0's build
            if exe.exists():
                break
            time.sleep(1)
    if not exe.exists():
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/bin/p2p_bench missing (run baseline/build_ref.sh; the reference is a C/CUDA library, not pip-installable)"}))
        return 0
    env = dict(os.environ)
    env["MPIACX_NFLAGS"] = str(max(4096, 4 * (a.steps + a.warmup) + 64))
    port = os.environ.get("MASTER_PORT", "0")
    outs = {}
    for phase, extra in (("dev", []), ("e2e", ["--e2e"])):
        env["ACX_JOBID"] = f"ref_{port}_{phase}_{os.environ.get('TORCHELASTIC_RUN_ID', 'x')}"
        cmd = [str(exe), "ring", "--bytes", str(a.bytes), "--steps", str(a.steps), "--warmup", str(max(3, a.warmup)), "--no-heap"] + extra
        from mpi_acx_b200.utils.clocks import ClockSampler
        # same placement policy as the acx arm (whose MPIX_Init binds itself): the rank runs on the socket its GPU hangs off,
        # so cudaMallocHost staging memory is first-touched there — what mpiexec's default binding would give the reference
        from mpi_acx_b200.utils.numa import gpu_numa_cpus
        cpus = gpu_numa_cpus(local) if os.environ.get("ACX_NUMA_BIND", "1") != "0" else None
        with ClockSampler(local, period_ms=50) as cs:
            p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1800, preexec_fn=(lambda: os.sched_setaffinity(0, cpus)) if cpus else None)
        if p.returncode != 0:
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": f"reference run failed rc={p.returncode}: {(p.stderr or p.stdout)[-300:]!r}"}))
            return 0
        if rank == 0:
            line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
            outs[phase] = json.loads(line)
            outs[phase + "_clocks"] = cs.summary
    extras = {} if a.no_extras else c_harness_extras(REPO / "baseline" / "_ref" / "bin", world, local, rank, True, "refx")
    if rank == 0:
        d, e = outs["dev"], outs["e2e"]
        print(json.dumps({
            "impl": "reference", "metric": "ring_sendrecv_bandwidth", "value": d["gbs_total"], "unit": "GB/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": d["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bytes",
            "data": "synthetic", "clocks": outs["dev_clocks"],
            "e2e": {"value": e["gbs_total"], "unit": "GB/s", "h2d_bytes_per_step": e["h2d_bytes_per_step"], "d2h_bytes_per_step": e["d2h_bytes_per_step"], "ms_per_step": e["ms_per_step"]},
            "gpu_launches": None,
            "config": {"model": "ring Isend/Irecv_enqueue + Waitall_enqueue", "msg_bytes": a.bytes, "global_batch": a.bytes * world, "seq_len": None,
                       "parallelism": f"p2p-ring{world}", "l2": "working set 2x msg_bytes > L2, no flush",
                       "order": "irecv, isend, waitall per step",
                       "note": "unmodified NVIDIA/mpi-acx (CPU proxy) over the bundled mini-MPI shim (shm + CUDA-IPC cudaMemcpyAsync); cudaMalloc buffers"},
            "extra": extras,
        }))
    return 0


# ------------------------------------------------------------------------------------------
def run_acx(a) -> int:
    import torch
    rank, world, local = env_rank()
    os.environ.setdefault("ACX_HEAP_BYTES", str(max(3 * a.bytes + (192 << 20), 512 << 20)))
    os.environ.setdefault("ACX_WATCHDOG_MS", "60000")
    import mpi_acx_b200 as acx
    from mpi_acx_b200.utils.clocks import ClockSampler
    # ACX_BENCH_DRYRUN_CPU=1: exercise this script's control flow on the host transport (no GPU, gloo, wall-clock
    # "events") — used by the CPU test-suite; its numbers mean nothing.
    dry = bool(os.environ.get("ACX_BENCH_DRYRUN_CPU"))
    dev = "cpu" if dry else "cuda"
    if dry:
        import time as _time

        class _Ev:
            def __init__(self, enable_timing=True): self.t = 0.0
            def record(self, stream=None): self.t = _time.perf_counter()
            def elapsed_time(self, other): return (other.t - self.t) * 1e3
        torch.cuda.Event = _Ev
        torch.cuda.synchronize = lambda *a_, **k_: None
        torch.cuda.current_stream = lambda *a_, **k_: None
        os.environ["ACX_FORCE_HOST"] = "1"
        a.no_extras = True
    elif not torch.cuda.is_available():
        print(json.dumps({"metric": "ring_sendrecv_bandwidth", "value": None, "error": "no CUDA device"}))
        return 1
    else:
        torch.cuda.set_device(local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        if dry:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    acx.init(local)
    assert acx.size() == world and acx.rank() == rank, (acx.size(), world)
    right, left = (rank + 1) % world, (rank - 1) % world
    n = a.bytes
    if a.order == "auto":
        a.order = "simple"   # both arms run the same per-step order (irecv, isend, waitall); prepost stays available as an explicit choice for both arms
    sbuf = acx.empty(n, torch.uint8)
    rbufs = [acx.empty(n, torch.uint8), acx.empty(n, torch.uint8)]
    rbuf = rbufs[0]
    sbuf.copy_(torch.arange(n, device=dev, dtype=torch.int32).to(torch.uint8) + rank)
    for b in rbufs:
        b.zero_()
    h_in = torch.empty(n, dtype=torch.uint8)
    h_out = torch.empty(4096, dtype=torch.uint8)
    if not dry:
        h_in, h_out = h_in.pin_memory(), h_out.pin_memory()   # after acx.init(): pages land on the GPU's NUMA node (MPIX_Init binds the rank)
    h_in.fill_(rank + 1)
    # a stream of our own, not the legacy default stream torch starts on: stream 0 synchronizes implicitly with every blocking
    # stream of the process and takes no programmatic dependent launches — the C harness (and any real application) enqueues
    # communication on an ordinary stream
    if not dry:
        torch.cuda.set_stream(torch.cuda.Stream())
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not use_dist:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # e2e: the step's input starts in pinned host memory. The library's host-staged exchange (parallel/staged.py) cuts the
    # message into chunks and sends chunk i (stream-triggered isend, gated on that chunk's H2D copy) while chunk i+1 is
    # still crossing PCIe; ACX_BENCH_E2E_CHUNKS=1 gives the plain copy-then-send sequence.
    from mpi_acx_b200.parallel.staged import StagedExchange
    e2e_chunks = int(os.environ.get("ACX_BENCH_E2E_CHUNKS", "8"))
    plan = StagedExchange(sbuf, rbuf, right, left, chunks=e2e_chunks) if (e2e_chunks > 1 and not dry) else None

    def step(e2e: bool):
        if e2e and plan is not None:
            plan.step(h_in)
            h_out.copy_(rbuf[:4096], non_blocking=True)
            return
        if e2e:
            sbuf.copy_(h_in, non_blocking=True)
        rr = acx.irecv(rbuf, left, 0)
        rs = acx.isend(sbuf, right, 0)
        acx.waitall_enqueue([rr, rs])
        if e2e:
            h_out.copy_(rbuf[:4096], non_blocking=True)

    from mpi_acx_b200.parallel import ring_exchange_prepost

    def run_steps(e2e: bool, k: int):
        """k ring steps; returns the buffer that holds the last received message."""
        if a.order == "simple":
            for _ in range(k):
                step(e2e)
            return rbuf
        before = (lambda i: sbuf.copy_(h_in, non_blocking=True)) if e2e else None
        after = (lambda i, buf: h_out.copy_(buf[:4096], non_blocking=True)) if e2e else None
        ring_exchange_prepost(sbuf, rbufs, k, 0, None, before, after)
        return rbufs[(k - 1) % 2]

    last_buf = {"t": rbuf}

    def timed(e2e: bool, steps: int, warmup: int):
        run_steps(e2e, warmup)
        barrier()
        k0 = acx.stats()["kernels_launched"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        last_buf["t"] = run_steps(e2e, steps)
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        return ms, acx.stats()["kernels_launched"] - k0

    # One nvidia-smi sampler (rank 0's GPU), started well before the first timed region and kept running across
    # both: starting NVML clients on every rank at the instant the timed loop begins stalls kernel launches.
    sampler = ClockSampler(local, period_ms=50) if rank == 0 and not dry else None
    if sampler:
        sampler.__enter__()
        import time
        time.sleep(0.5)
    barrier()

    W = max(3, a.warmup)
    ms, launches = timed(False, a.steps, W)
    # correctness of what was timed: the received buffer must be the left neighbour's pattern
    exp = (torch.arange(4096, device=dev, dtype=torch.int32).to(torch.uint8) + left)
    got = last_buf["t"]
    ok = bool(torch.equal(got[:4096], exp)) and bool(torch.equal(got[-4096:], (torch.arange(n - 4096, n, device=dev, dtype=torch.int64).to(torch.uint8) + left)))
    ms_e2e, _ = timed(True, a.steps, W)
    ok_e2e = bool((h_out == (left + 1)).all())
    clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    if sampler:
        sampler.__exit__(None, None, None)
        clocks = sampler.summary
        clocks["window"] = "warm-up + device-timed region + e2e region (the timed region alone is shorter than one sampling period)"
    gbs = world * n / ms / 1e6
    gbs_e2e = world * n / ms_e2e / 1e6

    e2e_desc = (f"{len(plan.bounds)} chunks: H2D of chunk i+1 overlaps the stream-triggered send of chunk i (mpi_acx_b200.parallel.staged)" if plan is not None else "copy, then send")
    extras = {}
    if not a.no_extras:
        # small-message ring step through the Python API: acx.sendrecv = MPIX_Sendrecv_enqueue, one call and (eager sizes)
        # one kernel per exchange step — receive posted, send deposited, receive completed
        for nb in (8, 4096, 1 << 20):
            s2, r2 = acx.empty(nb, torch.uint8), acx.empty(nb, torch.uint8)
            for _ in range(10):
                acx.sendrecv(s2, right, r2, left, 9)
            barrier()
            iters = 200
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                acx.sendrecv(s2, right, r2, left, 9)
            e1.record(stream)
            torch.cuda.synchronize()
            extras[f"ring_step_us_{nb}B_python"] = round(max_over_ranks(e0.elapsed_time(e1)) / iters * 1e3, 3)
            del s2, r2
        del sbuf, rbuf, got
        rbufs.clear()
        last_buf.clear()
        plan = None
        sbuf = rbuf = None
        barrier()
        # the same C harnesses the reference arm runs (identical keys), plus the fused kernels
        extras.update(c_harness_extras(REPO / "bin", world, local, rank, False, "acxx"))
        hm = extras.get("halo_fused_512cube_ms_per_step")
        if isinstance(hm, float):
            extras["halo_fused_hbm_frac_of_measured_6571GBs"] = round(2 * 2 * 512 ** 3 / hm / 1e6 / 6570.9, 3)
        gm = extras.get("ppgemm_fused_fp8_8192_ms_per_microbatch")
        if isinstance(gm, float):
            extras["ppgemm_fused_fp8_tflops_total"] = round(world * 2 * 8192 ** 3 / gm / 1e9, 1)
    out = {
        "metric": "ring_sendrecv_bandwidth", "value": round(gbs, 3), "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": W,
        "ms_per_step": round(ms, 6), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bytes", "data": "synthetic",
        "impl": "acx-b200", "verified": ok and ok_e2e, "clocks": clocks,
        "e2e": {"value": round(gbs_e2e, 3), "unit": "GB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": 4096, "ms_per_step": round(ms_e2e, 6),
                "pipeline": e2e_desc},
        "gpu_launches": int(launches),
        "config": {"model": "ring Isend/Irecv_enqueue + Waitall_enqueue", "msg_bytes": n, "global_batch": n * world, "seq_len": None,
                   "parallelism": f"p2p-ring{world}", "l2": "working set >= 2x msg_bytes > L2, no flush", "buffers": "symmetric heap (zero-copy NVLink push/pull)",
                   "order": "receive of step i+1 posted before step i is waited on (two receive buffers)" if a.order == "prepost" else "irecv, isend, waitall per step",
                   "nvlink_fraction_of_900GBs_per_direction_per_gpu": round(gbs / world / 900.0, 3) if world > 1 else None},
        "extra": extras,
    }
    if rank == 0:
        print(json.dumps(out))
        sys.stdout.flush()
    sbuf = rbuf = got = None
    rbufs.clear()
    last_buf.clear()
    barrier()
    acx.finalize()
    if use_dist:
        dist.destroy_process_group()
    return 0 if (ok and ok_e2e) else 1


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_acx(args))
