"""Engine copy bandwidth vs number of engine CTAs (one 64 MiB GPU-posted write stream, loopback over HBM)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W

ctx = rn.Context(0)
size = 64 << 20
src = torch.empty(4 * size, dtype=torch.uint8, device="cuda"); dst = torch.empty(4 * size, dtype=torch.uint8, device="cuda")
ops.fill_random(src, 3); torch.cuda.synchronize()
ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
qp = ctx.loopback_qp(depth=64)
rows = []
for ctas in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,2,4,8,16,32,64,96,128".split(","))]:
    ctx.engine_start(ctas=ctas, idle_timeout_ms=3000)
    kw = dict(window=8, slot_stride=size, nslots=4, timeout_ms=5000)
    ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=4, **kw)
    r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=24, **kw)
    ctx.engine_stop()
    rows.append(dict(engine_ctas=ctas, ok=r.ok, gbps=round(r.gbps, 1), gbps_per_cta=round(r.gbps / ctas, 1)))
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/engine_scaling.json", "w"), indent=1)
