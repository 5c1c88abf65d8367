"""Run-time configuration surface (env vars / CLI), the counterpart of the reference's build-time knobs
(KDIR, OFA_KERNEL_DIR, RDMA_HEADER_DIR: Makefile:2,23,46 -- it has no run-time configuration at all,
SURVEY.md section 5).  Every field can be set from ``ROCNRDMA_<NAME>`` in the environment."""
from __future__ import annotations

import os
from dataclasses import dataclass, field, fields
from typing import Dict, Optional

_SIZE_RE = {"k": 1 << 10, "m": 1 << 20, "g": 1 << 30}


def parse_size(s) -> int:
    """'64', '4k', '1M', '2GiB', '256 MiB' -> bytes."""
    if isinstance(s, int):
        return s
    t = str(s).strip().lower().replace("ib", "").replace("b", "").replace(" ", "")
    if not t:
        raise ValueError("empty size")
    mult = 1
    if t[-1] in _SIZE_RE:
        mult = _SIZE_RE[t[-1]]
        t = t[:-1]
    v = float(t)
    if v < 0 or v * mult != int(v * mult):
        raise ValueError(f"bad size {s!r}")
    return int(v * mult)


def parse_sweep(spec: str):
    """'1k:1g' -> powers of two from 1 KiB to 1 GiB; '1k:1m:x4' -> x4 steps; '4k,64k,1m' -> list."""
    spec = spec.strip()
    if "," in spec:
        return [parse_size(x) for x in spec.split(",") if x.strip()]
    parts = spec.split(":")
    lo = parse_size(parts[0])
    hi = parse_size(parts[1]) if len(parts) > 1 else lo
    step = 2
    if len(parts) > 2:
        step = int(parts[2].lstrip("x*"))
    if lo <= 0 or hi < lo or step < 2:
        raise ValueError(f"bad sweep {spec!r}")
    out, v = [], lo
    while v <= hi:
        out.append(v)
        v *= step
    return out


@dataclass
class Config:
    wire: str = "auto"              # auto | softhca | verbs
    registration: str = "auto"      # auto | direct | dmabuf | peermem | host-staged
    post: str = "gpu"               # gpu | host
    nic: str = ""                   # HCA name for the verbs wire ("" = the one closest to the GPU)
    port: int = 1
    gid_index: int = 0
    qp_depth: int = 256
    cq_depth: int = 512
    chunk_bytes: int = 512 << 10    # engine work granule
    engine_ctas: int = 32
    engine_idle_timeout_ms: int = 5000
    rnr_timeout_ms: int = 500
    queue_mem: str = "device"       # device | host  (where SQ/CQ rings live)
    affinity: Dict[int, str] = field(default_factory=dict)   # GPU index -> HCA name override

    @classmethod
    def from_env(cls, env: Optional[dict] = None) -> "Config":
        env = os.environ if env is None else env
        c = cls()
        for f in fields(cls):
            key = f"ROCNRDMA_{f.name.upper()}"
            if key not in env:
                continue
            raw = env[key]
            if f.name == "affinity":
                c.affinity = {int(k): v for k, v in (kv.split(":") for kv in raw.split(",") if kv)}
            elif f.name == "chunk_bytes":
                c.chunk_bytes = parse_size(raw)
            elif f.type in ("int", int):
                setattr(c, f.name, int(raw))
            else:
                setattr(c, f.name, raw)
        c.validate()
        return c

    def validate(self):
        if self.wire not in ("auto", "softhca", "verbs"):
            raise ValueError(f"wire={self.wire!r}")
        if self.registration not in ("auto", "direct", "dmabuf", "peermem", "host-staged"):
            raise ValueError(f"registration={self.registration!r}")
        if self.post not in ("gpu", "host"):
            raise ValueError(f"post={self.post!r}")
        if self.queue_mem not in ("device", "host"):
            raise ValueError(f"queue_mem={self.queue_mem!r}")
        for name in ("qp_depth", "cq_depth"):
            v = getattr(self, name)
            if v < 2 or v & (v - 1):
                raise ValueError(f"{name} must be a power of two >= 2")
        if self.chunk_bytes % 16 or self.chunk_bytes <= 0:
            raise ValueError("chunk_bytes must be a positive multiple of 16")
        return self

    def resolve(self, probe_result: dict) -> "Config":
        """Fill every 'auto' from a probe() result."""
        plan = probe_result["plan"]
        if self.wire == "auto":
            self.wire = plan["wire"] if plan["wire"] != "none" else "softhca"
        if self.registration == "auto":
            self.registration = plan["registration"] if plan["registration"] != "none" else "direct"
        return self
