#!/usr/bin/env python
"""b200p2ptest -- the userspace program the reference harness never shipped.

Drives the four harness verbs (+ bus-address readback and the CPU window) against either the kernel
module (/dev/b200p2ptest) or, where no module can be loaded, the CUDA-driver-API twin.

    b200p2ptest_cli.py selftest [--backend auto|dev|user] [--size 8m]
    b200p2ptest_cli.py classify <addr>
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocnrdma_b200 import harness as H  # noqa: E402
from rocnrdma_b200.config import parse_size  # noqa: E402

PAGE = H.GPU_PAGE_SIZE


def selftest(backend: str, size: int) -> dict:
    import torch
    be = H.open_backend(backend)
    kind = type(be).__name__
    t = torch.zeros(max(size, 4 << 20), dtype=torch.uint8, device="cuda:0")
    host = torch.zeros(4096, dtype=torch.uint8)
    va = t.data_ptr()
    res = {"backend": kind, "checks": {}}

    def check(name, ok, detail=""):
        res["checks"][name] = {"ok": bool(ok), "detail": detail}
        print(f"  [{'ok' if ok else 'FAIL'}] {name} {detail}")

    check("1 gpu address is classified as GPU", be.is_gpu_address(va + 17))
    check("1 host address is not", not be.is_gpu_address(host.data_ptr()))
    check("2 page size is 64 KiB", be.get_page_size(va, 2 * PAGE) == PAGE)
    g = be.get_pages(va, 4 * PAGE)
    check("3 pin returns one entry per GPU page", g.entries == 4 and g.page_size == PAGE, f"handle={g.handle}")
    if hasattr(be, "bus_addrs"):
        addrs = be.bus_addrs(g.handle)
        check("3 bus addresses are distinct and page aligned", len(set(addrs)) == 4 and all(a % 4096 == 0 for a in addrs),
              " ".join(hex(a) for a in addrs))
        m = be.mmap(va, 2 * PAGE)
        m[0:8] = b"BARwrite"
        torch.cuda.synchronize()
        check("3 CPU store through the BAR lands in HBM", bytes(t[:8].cpu().tolist()) == b"BARwrite")
        m.close()
    else:
        check("3 the kernel sizes the pinned object correctly", be.pin_size(g.handle) == 4 * PAGE)
        try:                                  # the harness's mmap, from userspace: a CPU mapping of the pin's dma-buf
            be.map_window(g.handle)
            res["cpu_window"] = be.window_kind(va)
        except H.HarnessError as e:
            res["cpu_window"] = f"{be.window_kind(va)}; mmap(dma-buf fd) failed: errno {e.errno} ({e.strerror})"
        print("  cpu window:", res["cpu_window"])
        be.poke(va, b"BARwrite")
        torch.cuda.synchronize()
        check("3 CPU window write lands in HBM", bytes(t[:8].cpu().tolist()) == b"BARwrite")
    be.get_pages(va, 4 * PAGE)
    be.get_pages(va, 4 * PAGE)
    n = be.put_pages(va, 4 * PAGE)
    check("4 one PUT_PAGES releases every pin of that exact range", n == 3, f"released={n}")
    be.get_pages(va, PAGE)
    be.get_pages(va + PAGE, PAGE)
    left = be.close()
    check("5 close releases what the application leaked", left in (2, None), f"leftover={left}")
    res["ok"] = all(c["ok"] for c in res["checks"].values())
    return res


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    st = sub.add_parser("selftest")
    st.add_argument("--backend", default="auto", choices=["auto", "dev", "user"])
    st.add_argument("--size", default="8m")
    st.add_argument("--json", default="")
    cl = sub.add_parser("classify")
    cl.add_argument("addr")
    cl.add_argument("--backend", default="auto")
    a = ap.parse_args(argv)
    if a.cmd == "selftest":
        r = selftest(a.backend, parse_size(a.size))
        if a.json:
            json.dump(r, open(a.json, "w"), indent=1)
        print("PASS" if r["ok"] else "FAIL", f"({r['backend']})")
        return 0 if r["ok"] else 1
    be = H.open_backend(a.backend)
    print(int(be.is_gpu_address(int(a.addr, 0))))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
