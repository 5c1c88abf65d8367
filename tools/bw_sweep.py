"""Engine-CTA sweep at large messages: GPU-posted RDMA write loopback, device-timed."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W
N = 1 << 30
ctx = rn.Context(0)
src = torch.empty(N, dtype=torch.uint8, device="cuda"); dst = torch.empty(N, dtype=torch.uint8, device="cuda")
ops.fill_random(src, 7); torch.cuda.synchronize()
ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
qp = ctx.loopback_qp(depth=256)
rows = []
for ctas in [int(a) for a in sys.argv[1:]] or [8, 32, 64, 128, 144]:
    ctx.engine_start(ctas=ctas, idle_timeout_ms=3000)
    for size in [1 << 20, 16 << 20, 256 << 20, 1 << 30]:
        iters = max(4, min(256, (4 << 30) // size)); nslots = max(1, N // size)
        ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=4, window=8, slot_stride=size, nslots=nslots)
        r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=iters, window=8, slot_stride=size, nslots=nslots)
        rows.append(dict(ctas=ctas, size=size, ok=r.ok, gbps=round(r.gbps, 1)))
        print(rows[-1], flush=True)
    ctx.engine_stop()
print("verify", ops.compare(src, dst) == 0)
json.dump(rows, open("gpurun_out/bw_sweep.json", "w"))
