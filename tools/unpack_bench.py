"""K5 alone: fp8 chunk records -> bf16, device-timed, against the algorithmic bytes at the measured copy peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops

ctx = rn.Context(0, wire="softhca")
for n, chunk in ((1 << 28, 1 << 22), (1 << 28, 1 << 20), (1 << 26, 1 << 22)):
    x = torch.randn(n, device="cuda").to(torch.bfloat16)
    nb = ops.staging_bytes(n, chunk)
    stg = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ops.pack_fp8_write(ctx, x, ctx.reg_mr(stg), chunk_elems=chunk)
    y = torch.zeros_like(x)
    best = min(ops.unpack_fp8(ctx, stg, y, chunk_elems=chunk)["device_ns"] for _ in range(5))
    ok = torch.equal(y, ops.ref_unpack_fp8(stg, n, chunk))
    print(f"n={n >> 20} Mi chunk={chunk >> 20} Mi: {best / 1e3:.1f} us, {(nb + 2 * n) / best:.0f} GB/s of HBM traffic (copy peak ~6580), ok={ok}", flush=True)
ctx.close()
