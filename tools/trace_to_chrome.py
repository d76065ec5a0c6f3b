#!/usr/bin/env python3
"""Merge the per-rank device traces (ACX_TRACE=1 -> acx_trace_rank<r>.jsonl, one %globaltimer-stamped record per protocol
event) into one Chrome / Perfetto trace (chrome://tracing, ui.perfetto.dev):

    ACX_TRACE=1 ACX_TRACE_DIR=/tmp/t bin/acxrun -np 2 build/tests/ring_all_device
    tools/trace_to_chrome.py /tmp/t -o trace.json

Every rank is a process row; waits and data movement become duration slices (wait_begin..wait_end, move_begin..move_end,
matched per peer and ticket), triggers become instant events, and each message gets a flow arrow from the sender's
send_trigger to the receiver's wait_end with the same ticket. The reference has no tracing at all.
Also prints a per-rank summary (events, time in waits, bytes moved). No GPU needed to run it."""
import argparse
import glob
import json
import os
import sys


def load(path_or_dir):
    files = sorted(glob.glob(os.path.join(path_or_dir, "acx_trace_rank*.jsonl"))) if os.path.isdir(path_or_dir) else [path_or_dir]
    recs = []
    for f in files:
        with open(f) as fh:
            for line in fh:
                line = line.strip()
                if line:
                    recs.append(json.loads(line))
    return recs


def convert(recs):
    if not recs:
        return {"traceEvents": []}, {}
    t0 = min(r["t_ns"] for r in recs)
    us = lambda r: (r["t_ns"] - t0) / 1000.0
    ev = []
    summary = {}
    open_slices = {}   # (rank, kind, peer, ticket) -> begin record
    flow_id = 0
    sends = {}         # (sender, receiver, ticket) -> send_trigger record
    tid_of = {"send_trigger": 1, "recv_trigger": 2, "move": 3, "wait": 4}
    for rank in sorted({r["rank"] for r in recs}):
        ev.append({"name": "process_name", "ph": "M", "pid": rank, "args": {"name": f"rank {rank}"}})
        for name, tid in (("send triggers", 1), ("recv triggers", 2), ("movers", 3), ("waits", 4)):
            ev.append({"name": "thread_name", "ph": "M", "pid": rank, "tid": tid, "args": {"name": name}})
        summary[rank] = {"events": 0, "wait_us": 0.0, "move_us": 0.0, "bytes_moved": 0, "sends": 0, "recvs": 0}
    for r in sorted(recs, key=lambda x: x["t_ns"]):
        rank, e, peer, a, b = r["rank"], r["ev"], r["peer"], r["a"], r["b"]
        s = summary[rank]
        s["events"] += 1
        if e in ("send_trigger", "recv_trigger"):
            rndv = bool(b >> 63) if e == "send_trigger" else False
            nbytes = b & ((1 << 63) - 1)
            ev.append({"name": e, "ph": "i", "s": "t", "pid": rank, "tid": tid_of[e], "ts": us(r),
                       "args": {"peer": peer, "ticket": a, "bytes": nbytes, "rendezvous": rndv}})
            s["sends" if e == "send_trigger" else "recvs"] += 1
            if e == "send_trigger":
                sends[(rank, peer, a)] = r
        elif e in ("wait_begin", "move_begin"):
            open_slices[(rank, e[:4], peer, a)] = r
        elif e in ("wait_end", "move_end"):
            kind = e[:4]
            beg = open_slices.pop((rank, kind, peer, a), None)
            if beg is None:
                continue
            dur = us(r) - us(beg)
            ev.append({"name": f"{kind} peer {peer} #{a}", "ph": "X", "pid": rank, "tid": tid_of[kind], "ts": us(beg), "dur": max(dur, 0.001),
                       "args": {"peer": peer, "ticket": a, "bytes": b}})
            if kind == "wait":
                s["wait_us"] += dur
                snd = sends.get((peer, rank, a))   # the message this receive-wait completed (waits on sends find nothing: fine)
                if snd is not None and beg["b"] == 1:
                    flow_id += 1
                    ev.append({"name": "msg", "cat": "msg", "ph": "s", "id": flow_id, "pid": peer, "tid": 1, "ts": us(snd)})
                    ev.append({"name": "msg", "cat": "msg", "ph": "f", "bp": "e", "id": flow_id, "pid": rank, "tid": 4, "ts": us(r)})
            else:
                s["move_us"] += dur
                s["bytes_moved"] += b
    return {"traceEvents": ev, "displayTimeUnit": "ns"}, summary


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("input", help="directory holding acx_trace_rank*.jsonl, or one such file")
    ap.add_argument("-o", "--output", default="acx_trace.json")
    a = ap.parse_args()
    recs = load(a.input)
    if not recs:
        print("no trace records found", file=sys.stderr)
        return 1
    trace, summary = convert(recs)
    with open(a.output, "w") as fh:
        json.dump(trace, fh)
    for rank, s in sorted(summary.items()):
        gbps = s["bytes_moved"] / (s["move_us"] * 1e3) if s["move_us"] > 0 else 0.0
        print(f"rank {rank}: {s['events']} events, {s['sends']} sends, {s['recvs']} recvs, {s['wait_us']:.1f} us in waits, "
              f"{s['bytes_moved']} B moved in {s['move_us']:.1f} us ({gbps:.1f} GB/s while moving)")
    print(f"wrote {a.output}: {len(trace['traceEvents'])} events")
    return 0


if __name__ == "__main__":
    sys.exit(main())
