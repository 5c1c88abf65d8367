"""First-light sweep: GPU-posted RDMA write loopback, device-timed, a few sizes."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W

ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 64
import time
T0=time.time()
def log(*a):
    print(f"[{time.time()-T0:6.2f}]", *a, flush=True)
    open("gpurun_out/first_light_progress.txt","a").write(" ".join(map(str,a))+"\n")
os.makedirs("gpurun_out", exist_ok=True)
log("start", ctas)
ctx = rn.Context(0)
N = 1 << 30
src = torch.empty(N, dtype=torch.uint8, device="cuda"); dst = torch.empty(N, dtype=torch.uint8, device="cuda")
ops.fill_random(src, 7); torch.cuda.synchronize(); log("filled")
ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
qp = ctx.loopback_qp(depth=256)
log("qp"); ctx.engine_start(ctas=ctas, idle_timeout_ms=4000); log("engine")
rows = []
for size in [64, 1024, 16384, 262144, 4 << 20, 64 << 20, 1 << 30]:
    iters = max(4, min(2000, (2 << 30) // size))
    nslots = max(1, min(N // size, 4096))
    log('size', size, iters, nslots)
    w = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=min(iters, 50), window=32, slot_stride=size, nslots=nslots)
    log('warm', w.status, w.done, ctx.engine_stats())
    r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=iters, window=32, slot_stride=size, nslots=nslots)
    r1 = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=min(iters, 200), window=1, slot_stride=size, nslots=nslots)
    rows.append(dict(size=size, iters=iters, ok=r.ok, gbps=round(r.gbps, 2), us_per_msg=round(r.us_per_msg, 2), lat_us=round(r1.us_per_msg, 2)))
    print(rows[-1], flush=True)
print(ctx.engine_stats(), qp.counters())
ctx.engine_stop()
ok = ops.compare(src, dst) == 0
print("verify", ok)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(ctas=ctas, rows=rows, verify=ok), open(f"gpurun_out/first_light_{ctas}.json", "w"))
