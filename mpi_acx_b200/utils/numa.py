"""NUMA helpers for launch scripts: which CPUs sit on the socket a GPU hangs off (the C runtime binds itself in
MPIX_Init, see acx_numa_bind in src/core/acx.cu; this is for harnesses that start other programs)."""
from __future__ import annotations

import os
import subprocess
from typing import Optional, Set


def _parse_cpulist(text: str) -> Set[int]:
    cpus: Set[int] = set()
    for tok in text.strip().split(","):
        if not tok:
            continue
        if "-" in tok:
            a, b = tok.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(tok))
    return cpus


def gpu_numa_cpus(gpu_index: int) -> Optional[Set[int]]:
    """CPUs of the NUMA node of GPU `gpu_index` (None when it cannot be determined)."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)], capture_output=True, text=True, timeout=20).stdout.strip()
        bus = out.splitlines()[0].strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:   # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_socket(gpu_index: int) -> bool:
    """Restrict the calling process to the CPUs of the GPU's socket (first-touch then places its memory there)."""
    cpus = gpu_numa_cpus(gpu_index)
    if not cpus:
        return False
    os.sched_setaffinity(0, cpus)
    return True
