"""Sustained GEMM loops (about 2 s each) with the SM clock and board power sampled meanwhile: ours vs cuBLAS.
python tools/gemm_power.py 8192x8192x8192 [...]   (RN_CTA_GROUP selects the kernel)"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops

shapes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[1:] or ["8192x8192x8192"])]
ctx = rn.Context(0, wire="softhca")
kw = {}
if os.environ.get("RN_CTA_GROUP"):
    kw["cta_group"] = int(os.environ["RN_CTA_GROUP"])


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop_flag = [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((time.time(), float(out[0]), float(out[1]), float(out[2])))
            except Exception:
                pass
            time.sleep(0.05)


def med(v):
    v = sorted(v)
    return v[len(v) // 2] if v else 0.0


for (M, N, K) in shapes:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    c = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    flops = 2.0 * M * N * K
    for name, fn in (("ours", lambda: ops.gemm_send(ctx, a, b, c, sync=False, stream=ctx.stream, **kw)), ("cublas", lambda: torch.matmul(a, b.T, out=c)),
                     ("ours", lambda: ops.gemm_send(ctx, a, b, c, sync=False, stream=ctx.stream, **kw)), ("cublas", lambda: torch.matmul(a, b.T, out=c))):
        n = max(20, int(2.0 / (flops / 1.5e15)))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with torch.cuda.stream(ctx.stream):
            for _ in range(5):
                fn()
        ctx.stream.synchronize()
        sm = Sampler(); sm.start()
        t0 = time.time()
        with torch.cuda.stream(ctx.stream):
            ev[0].record()
            for _ in range(n):
                fn()
            ev[1].record()
        ev[1].synchronize()
        sm.stop_flag = True; sm.join()
        rows = [r for r in sm.rows if r[0] > t0 + 0.5]
        tf = flops * n / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e12
        print(f"{M}x{N}x{K} {name:7s} {tf:7.1f} TF  sm {med([r[1] for r in rows]):.0f} MHz  power {med([r[2] for r in rows]):.0f} W  temp {med([r[3] for r in rows]):.0f} C  ({len(rows)} samples, {n} launches)", flush=True)
        time.sleep(1.0)
ctx.close()
