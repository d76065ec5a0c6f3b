#!/usr/bin/env python3
"""Randomized option-combination runs of the conformance programs on the software CUDA device (make emu-lib emu-tests):
    tools/emu_fuzz.py [seed] [seconds]
Each run picks a program, 1-4 ranks and 1-4 option groups (cooperative moving, per-peer branches, sender push, batched /
memOp partitioned rounds, fusion and side streams off, mover variants, tiny slot tables, arrival jitter, emulated device
shapes, ...). Combinations a program cannot support by design are skipped. Two real bugs came out of the first sweeps
(findings 6 and 7); the last sweeps (4300 runs) were clean.
EMU_DIR=build/emu_tsan EMU_FUZZ_NP1=1 runs the ThreadSanitizer build with both ends of every message in one process."""
import os, random, subprocess, sys, time
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600
opts = [("ACX_COOP=1", "ACX_COOP_MIN_BYTES=65536"), ("ACX_PEER_BRANCH=1",), ("ACX_SENDER_PUSH=1",), ("ACX_PBATCH=1",), ("ACX_MEMOPS=1",), ("ACX_FUSE_RNDV=0",),
        ("ACX_SIDE_STREAMS=0",), ("ACX_FUSE_MIN_CTAS=1", "ACX_EMU_SMS=16"), ("ACX_MOVER=ldst256",), ("ACX_MOVER_BLOCK=4096",), ("ACX_REGCACHE=0",),
        ("ACX_FAULT_DELAY_US=200",), ("ACX_NFLAGS=16",), ("ACX_EMU_MAX_BLOCK=32",), ("ACX_EMU_NO_HOSTREG_UVA=1",), ("ACX_PDL=0",), ("ACX_EMU_SMS=2",), ("ACX_EMU_JITTER_US=300",), ("ACX_EMU_JITTER_US=2000",), ("ACX_WILDCARDS=1",), ("ACX_EAGER_MAX=512",), ("ACX_MOVER_CTA_BYTES=4096",), ("ACX_MOVER_MAX_CTAS=3",),
        # round 2: sender-moves protocol and copy-engine rendezvous at emulation-sized messages, their variants, exact-tag directory off, hardware work queues
        ("ACX_LATE_PUSH_MIN=65536",), ("ACX_LATE_PUSH_MIN=65536", "ACX_LATE_FALLBACK_US=0"), ("ACX_CE_MIN=16384", "ACX_LATE_PUSH_MIN=65536"), ("ACX_CE_MIN=16384", "ACX_CE_SPINGATE=0"),
        ("ACX_CE_MIN=16384", "ACX_CE_MEMOP=0"), ("ACX_CE_MIN=16384", "ACX_CE_MEMOP=2", "ACX_CE_SPINGATE=0"), ("ACX_CE_MIN=16384", "ACX_CE_SPIN_US=0"), ("ACX_TAGDIR=0",),
        ("ACX_EMU_HW_QUEUES=1",), ("ACX_EMU_HW_QUEUES=2",), ("ACX_EMU_HW_QUEUES=3",),
        # overflow sends: a ring of 8 turns the bursts of many_outstanding / status_errors / stress into overflow sends; off = the old behaviour
        ("ACX_DEPTH=8",), ("ACX_DEPTH=8", "ACX_EMU_HW_QUEUES=1"), ("ACX_SEND_OVERFLOW=0",)]
progs_extra = ["wildcards", "tags", "many_outstanding", "host_reuse", "reinit", "late_push", "ce_path"]
sized = {"late_push": "262144", "ce_path": "262144"}   # programs that take their message size as an argument
progs = "ring ring_all ring_all_device ring_all_graph ring_graph_construction graph_large ring_partitioned partitioned_device ordering large_msgs regcache procnull_comms subcomm datatypes threads status_errors sendrecv_enqueue stress".split()
t0 = time.time(); runs = fails = 0
while time.time() - t0 < budget:
    k = random.choice([1, 2, 2, 3, 3, 4])
    chosen = random.sample(opts, k)
    env = dict(os.environ, ACX_MPI_NO_CUDA="1", ACX_WATCHDOG_MS="40000")
    for grp in chosen:
        for kv in grp:
            a, b = kv.split("="); env[a] = b
    prog = random.choice(progs + progs_extra); np_ = random.choice([1, 2, 3, 4]) if not os.environ.get("EMU_FUZZ_NP1") else 1   # EMU_FUZZ_NP1=1: both ends of every message in one process (ThreadSanitizer builds)
    if "ACX_NFLAGS" in env and prog in ("large_msgs", "status_errors", "graph_large"): continue
    if env.get("ACX_REGCACHE") == "0" and prog == "regcache": continue
    if "ACX_WILDCARDS" in env and prog == "procnull_comms": continue
    if prog == "wildcards": env["ACX_WILDCARDS"] = "1"
    if "ACX_EAGER_MAX" in env and prog in ("sendrecv_enqueue", "status_errors"): continue   # these assume the default eager limit
    if "ACX_EMU_HW_QUEUES" in env and prog == "threads" and np_ > 1: continue   # several application threads per rank need a hardware queue per stream
    if prog == "ce_path" and ("ACX_CE_MIN" not in env or any(k in env for k in ("ACX_COOP", "ACX_SENDER_PUSH", "ACX_PEER_BRANCH", "ACX_SIDE_STREAMS", "ACX_WILDCARDS", "ACX_MOVER_MAX_CTAS", "ACX_NFLAGS"))): continue   # asserts that the copy engines were used
    if "ACX_TAGDIR" in env and prog in ("tags", "threads"): continue   # need one channel per tag
    if "ACX_NFLAGS" in env and prog in ("many_outstanding", "tags", "host_reuse"): continue     # more than 16 requests at a time
    if "ACX_DEPTH" in env and ("ACX_SENDER_PUSH" in env or "ACX_PEER_BRANCH" in env): continue      # no overflow path in these modes: bursts longer than the ring stop their stream
    if "ACX_EMU_HW_QUEUES" in env and "ACX_SENDER_PUSH" in env: continue   # the push gate waits on a channel stream with its mover queued behind it: needs a hardware queue of its own (opt-in mode)
    if env.get("ACX_SEND_OVERFLOW") == "0" and "ACX_DEPTH" in env: continue   # bursts longer than the ring need the overflow path
    r = subprocess.run(["bin/acxrun", "-np", str(np_), "--timeout", "240", os.environ.get("EMU_DIR", "build/emu") + f"/tests/{prog}"] + ([sized[prog]] if prog in sized else []), env=env, capture_output=True, text=True, timeout=400)
    runs += 1
    if r.returncode != 0 or "PASS" not in r.stdout or "ThreadSanitizer" in r.stderr or "AddressSanitizer" in r.stderr or "runtime error:" in r.stderr:
        fails += 1
        print("FAIL", prog, "np", np_, " ".join(kv for g in chosen for kv in g), "|", (r.stdout + r.stderr).strip().splitlines()[-3:], flush=True)
print(f"fuzz: {runs} runs, {fails} failures")
