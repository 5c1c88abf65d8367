"""Build kmod/libb200p2p_sim.so: both kernel modules compiled unchanged against the userspace shim."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KMOD = ROOT / "kmod"
LIB = KMOD / "libb200p2p_sim.so"
FLAGS = ["-O1", "-g", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror=implicit-function-declaration",
         "-DB200P2P_SIM", f"-I{KMOD / 'shim'}", f"-I{KMOD / 'include'}"]


def build(force: bool = False) -> Path:
    """RN_KMOD_SIM_SANITIZE=1 builds a SEPARATE library with UBSan (=asan: + AddressSanitizer, the caller preloads
    libasan): the module sources run their whole simulated life cycle under the sanitizers
    (`make -C kmod check-sanitize`).  The plain library is never replaced by an instrumented one."""
    flags, link, lib, tag = list(FLAGS), [], LIB, "sim"
    mode = os.environ.get("RN_KMOD_SIM_SANITIZE")
    if mode:
        san = ["-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
        if mode == "asan":
            san.insert(0, "-fsanitize=address")
        flags += san
        link += san
        tag = "san_asan" if mode == "asan" else "san"
        lib = KMOD / f"libb200p2p_{tag}.so"
    srcs = [(KMOD / "b200p2p.c", "b200p2p"), (KMOD / "b200p2ptest.c", "b200p2ptest"), (KMOD / "shim" / "sim_runtime.c", "sim")]
    deps = [s for s, _ in srcs] + list((KMOD / "shim").rglob("*.h")) + list((KMOD / "include").glob("*.h"))
    if not force and lib.exists() and all(lib.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return lib
    objs = []
    for src, mod in srcs:
        obj = KMOD / f".{mod}.{tag}.o"
        subprocess.run(["gcc", *flags, f"-DKBUILD_MODNAME={mod}", "-c", str(src), "-o", str(obj)], check=True)
        objs.append(str(obj))
    subprocess.run(["gcc", "-shared", "-o", str(lib), *objs, *link, "-lpthread"], check=True)
    return lib


def build_stress(sanitizer: str = "address") -> Path:
    """kmod/tests/sim_stress.c + both modules + the simulation runtime as ONE executable under a sanitizer
    ("address" or "thread"; TSan needs the whole program instrumented, so this is not the shared library)."""
    out = KMOD / f"sim_stress_{sanitizer}"
    srcs = [(KMOD / "b200p2p.c", "b200p2p"), (KMOD / "b200p2ptest.c", "b200p2ptest"), (KMOD / "shim" / "sim_runtime.c", "sim"),
            (KMOD / "tests" / "sim_stress.c", "stress")]
    deps = [s for s, _ in srcs] + list((KMOD / "shim").rglob("*.h")) + list((KMOD / "include").glob("*.h"))
    if out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return out
    san = [f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer"]
    flags = [f for f in FLAGS if f != "-fvisibility=hidden"] + san
    objs = []
    for src, mod in srcs:
        obj = KMOD / f".{mod}.stress_{sanitizer}.o"
        subprocess.run(["gcc", *flags, f"-DKBUILD_MODNAME={mod}", "-c", str(src), "-o", str(obj)], check=True)
        objs.append(str(obj))
    subprocess.run(["gcc", "-o", str(out), *objs, *san, "-lpthread"], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
