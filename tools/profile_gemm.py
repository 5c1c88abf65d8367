"""ncu target: our tcgen05 GEMM and cuBLAS (torch.matmul) back to back on the same operands, compute only.
Run:  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/gemm_vs_cublas python tools/profile_gemm.py
Without ncu it prints event-timed TFLOP/s of both (10 back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops

shapes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[1:] or ["8192x8192x8192", "4096x4096x4096"])]
ctx = rn.Context(0, wire="softhca")
bufs = {}
for (M, N, K) in shapes:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bufs[(M, N, K)] = (a, b, torch.zeros(M, N, device="cuda", dtype=torch.bfloat16))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
kw = {}
if os.environ.get("RN_GROUP_M"):
    kw["group_m"] = int(os.environ["RN_GROUP_M"])
if os.environ.get("RN_CTA_GROUP"):
    kw["cta_group"] = int(os.environ["RN_CTA_GROUP"])
for (M, N, K), (a, b, c) in bufs.items():
    flops = 2.0 * M * N * K
    res = {}
    for name, fn in (("ours", lambda: ops.gemm_send(ctx, a, b, c, sync=False, stream=ctx.stream, **kw)), ("cublas", lambda: torch.matmul(a, b.T, out=c))):
        with torch.cuda.stream(ctx.stream):
            for _ in range(3):
                fn()
            ev[0].record()
            for _ in range(10):
                fn()
            ev[1].record()
        ev[1].synchronize()
        res[name] = flops * 10 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e12
    single = max(ops.gemm_send(ctx, a, b, c, **kw).tflops for _ in range(5))
    print(f"{M}x{N}x{K}: ours {res['ours']:.1f} (in-kernel best {single:.1f})  cublas {res['cublas']:.1f}  ratio {res['ours'] / res['cublas']:.3f}", flush=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for (M, N, K), (a, b, c) in bufs.items():
    with torch.cuda.stream(ctx.stream):
        ops.gemm_send(ctx, a, b, c, sync=False, stream=ctx.stream, **kw)
        torch.matmul(a, b.T, out=c)
    ctx.stream.synchronize()
torch.cuda.profiler.stop()
ctx.close()
