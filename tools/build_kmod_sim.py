"""Build kmod/libb200p2p_sim.so: both kernel modules compiled unchanged against the userspace shim."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KMOD = ROOT / "kmod"
LIB = KMOD / "libb200p2p_sim.so"
FLAGS = ["-O1", "-g", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror=implicit-function-declaration",
         "-DB200P2P_SIM", f"-I{KMOD / 'shim'}", f"-I{KMOD / 'include'}"]


def build(force: bool = False) -> Path:
    srcs = [(KMOD / "b200p2p.c", "b200p2p"), (KMOD / "b200p2ptest.c", "b200p2ptest"), (KMOD / "shim" / "sim_runtime.c", "sim")]
    deps = [s for s, _ in srcs] + list((KMOD / "shim").rglob("*.h")) + list((KMOD / "include").glob("*.h"))
    if not force and LIB.exists() and all(LIB.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return LIB
    objs = []
    for src, mod in srcs:
        obj = KMOD / f".{mod}.sim.o"
        subprocess.run(["gcc", *FLAGS, f"-DKBUILD_MODNAME={mod}", "-c", str(src), "-o", str(obj)], check=True)
        objs.append(str(obj))
    subprocess.run(["gcc", "-shared", "-o", str(LIB), *objs, "-lpthread"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
