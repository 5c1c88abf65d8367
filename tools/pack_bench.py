"""K3 bench: fused bf16->fp8 pack + GPU-initiated RDMA write vs unfused (pack kernel, then host-synchronised write)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W
from rocnrdma_b200.ops import pack as P
ctx = rn.Context(0)
N = 1 << 29     # 512 Mi elements = 1 GiB bf16
x = torch.empty(N, dtype=torch.bfloat16, device="cuda"); ops.fill_bf16(x, 3, 1.0)
rows = []
for chunk in [1 << 20, 1 << 22]:
    nb = P.staging_bytes(N, chunk)
    staging = torch.empty(nb, dtype=torch.uint8, device="cuda"); remote = torch.empty(nb, dtype=torch.uint8, device="cuda")
    smr, rmr = ctx.reg_mr(staging), ctx.reg_mr(remote)
    qp = ctx.loopback_qp(depth=1024, cq_depth=2048)
    torch.cuda.synchronize()
    for ectas in [32, 64]:
        ctx.engine_start(ctas=ectas, idle_timeout_ms=3000)
        for sizeN in [1 << 24, 1 << 27, 1 << 29]:
            xs = x[:sizeN]
            P.pack_fp8_write(ctx, xs, smr, qp=qp, dst_mr=rmr, chunk_elems=chunk, signal_every=8)
            r = P.pack_fp8_write(ctx, xs, smr, qp=qp, dst_mr=rmr, chunk_elems=chunk, signal_every=8)
            # unfused: pack only, then one GPU-posted write of the whole staging region (two launches, host sync between)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            with torch.cuda.stream(ctx.stream):
                ev[0].record()
                P.pack_fp8_write(ctx, xs, smr, qp=None, chunk_elems=chunk, sync=True)
                wr = ops.rdma_stream(qp, W.OP_RDMA_WRITE, smr, rmr, min(P.staging_bytes(sizeN, chunk), 1 << 30), iters=1)
                ev[1].record(); ev[1].synchronize()
            unf_ms = ev[0].elapsed_time(ev[1])
            rows.append(dict(chunk=chunk, engine_ctas=ectas, elems=sizeN, ok=r.ok, fused_us=round(r.device_ns / 1e3, 1),
                             fused_wire_gbps=round(r.payload_gbps, 1), fused_src_gbps=round(r.source_gbps, 1),
                             pack_phase_us=round((r.t_pack_end_ns - r.t_start_ns) / 1e3, 1), unfused_us=round(unf_ms * 1e3, 1)))
            print(rows[-1], flush=True)
        ctx.engine_stop()
    ok = torch.equal(remote[:P.staging_bytes(1 << 24, chunk)], P.ref_pack_fp8(x[:1 << 24], chunk))
    print("verify", ok)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/pack_bench.json", "w"), indent=1)
