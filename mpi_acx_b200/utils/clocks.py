"""Sample SM clocks / throttle reasons with nvidia-smi while a timed region runs
(B200_PROFILING.md "clocks line"). Used by bench.py; harmless when nvidia-smi is absent."""
from __future__ import annotations

import statistics
import subprocess
from typing import Optional

_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")


class ClockSampler:
    def __init__(self, gpu_index: int = 0, period_ms: int = 100):
        self.gpu, self.period, self.p = gpu_index, period_ms, None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={_Q}", "--format=csv,noheader,nounits", "-lms", str(self.period), "-i", str(self.gpu)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None
        return self

    def __exit__(self, *a):
        self.summary = self._finish()

    def _finish(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons, busy = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                if len(f) > 8 and float(f[8]) >= 10.0:
                    busy.append(float(f[1]))
            except ValueError:
                pass
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        load = busy or sm   # median over samples taken while the GPU was busy; all samples if none qualified
        return {"sm_mhz": statistics.median(load) if load else None, "samples_under_load": len(busy), "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}
