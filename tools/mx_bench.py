"""K7 throughput: block-scaled fp8 GEMM vs our bf16 GEMM vs cuBLAS bf16 on the same logical shapes (event-timed, 10 launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops
from rocnrdma_b200.ops import gemm_mx as MX
ctx = rn.Context(0, wire="softhca")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (8192, 8192, 2048)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    c = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    (aq, as_), (bq, bs) = MX.quantize_mx(a), MX.quantize_mx(b)
    oa, ob = MX.MxOperand.from_tensors(aq, as_), MX.MxOperand.from_tensors(bq, bs)
    flops = 2.0 * M * N * K
    res = {}
    for name, fn in (("mxfp8", lambda: ops.gemm_mxfp8(ctx, oa, ob, c, sync=False, stream=ctx.stream)),
                     ("bf16_ours", lambda: ops.gemm_send(ctx, a, b, c, sync=False, stream=ctx.stream)),
                     ("bf16_cublas", lambda: torch.matmul(a, b.T, out=c))):
        with torch.cuda.stream(ctx.stream):
            for _ in range(3): fn()
            ev[0].record()
            for _ in range(10): fn()
            ev[1].record()
        ev[1].synchronize()
        res[name] = round(flops * 10 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e12, 1)
    print(f"{M}x{N}x{K}", res, flush=True)
ctx.close()
