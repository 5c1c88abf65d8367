"""Sample SM clocks / throttle reasons with nvidia-smi while a timed region runs."""
from __future__ import annotations

import os
import statistics
import subprocess
import threading
import time
from typing import List, Optional

_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


def visible_gpu_index(local_index: int):
    """nvidia-smi's index (or UUID) of CUDA device ``local_index`` of this process: CUDA_VISIBLE_DEVICES renumbers
    devices, nvidia-smi does not.  The sampler must watch the GPU under load -- polling every GPU of the node
    makes the median the clock of the idle ones."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if local_index < len(ids):
            v = ids[local_index]
            return int(v) if v.isdigit() else v          # an index, or a GPU-/MIG- UUID (nvidia-smi -i takes both)
    return local_index


class ClockSampler:
    """Background poller (one nvidia-smi call per period; cheap, and off the GPU's critical path)."""

    def __init__(self, gpu_index: Optional[int] = None, period_s: float = 0.1):
        self.gpu_index, self.period_s = gpu_index, period_s
        self.rows: List[dict] = []
        self._stop = threading.Event()
        self._thr: Optional[threading.Thread] = None

    def _poll_once(self):
        cmd = ["nvidia-smi", f"--query-gpu={_QUERY}", "--format=csv,noheader,nounits"]
        if self.gpu_index is not None:
            cmd += ["-i", str(self.gpu_index)]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=5).stdout
        except Exception:
            return
        now = time.monotonic()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                self.rows.append(dict(t=now, index=int(f[0]), sm=float(f[1]), sm_max=float(f[2]), power=float(f[3]),
                                      active=f[4], hw_slowdown=f[5], hw_thermal=f[6], sw_thermal=f[7], sw_power=f[8]))
            except ValueError:
                continue

    def _run(self):
        while not self._stop.is_set():
            self._poll_once()
            self._stop.wait(self.period_s)

    def start(self):
        self._t0 = time.monotonic()
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def stop(self, skip_first_s: float = 0.0) -> dict:
        """``skip_first_s`` drops the samples of the ramp (clocks rise over the first few hundred ms of load)
        as long as later ones exist."""
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=6)
        if skip_first_s > 0:
            late = [r for r in self.rows if r["t"] - self._t0 >= skip_first_s]
            if late:
                self.rows = late
        return self.summary()

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for r in self.rows:
            for key, name in (("hw_slowdown", "hw_slowdown"), ("hw_thermal", "hw_thermal_slowdown"),
                              ("sw_thermal", "sw_thermal_slowdown"), ("sw_power", "sw_power_cap")):
                if r[key].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r["sm"] for r in self.rows), "sm_max_mhz": max(r["sm_max"] for r in self.rows),
                "power_w_max": max(r["power"] for r in self.rows), "reasons": sorted(reasons), "samples": len(self.rows)}
