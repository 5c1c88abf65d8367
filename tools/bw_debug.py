import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
T0 = time.time()
os.makedirs("gpurun_out", exist_ok=True)
def log(*a):
    line = f"[{time.time()-T0:7.2f}] " + " ".join(map(str, a))
    print(line, flush=True); open("gpurun_out/bw_debug.txt", "a").write(line + "\n")
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W
N = 256 << 20
ctx = rn.Context(0)
src = torch.empty(N, dtype=torch.uint8, device="cuda"); dst = torch.empty(N, dtype=torch.uint8, device="cuda")
ops.fill_random(src, 7); torch.cuda.synchronize(); log("filled")
ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
qp = ctx.loopback_qp(depth=256)
ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx.engine_start(ctas=ctas, idle_timeout_ms=2000); log("engine", ctas)
for size, iters, window in [(1 << 20, 4, 1), (1 << 20, 4, 2), (1 << 20, 16, 8), (16 << 20, 8, 4), (256 << 20, 4, 2)]:
    nslots = max(1, N // size)
    r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=iters, window=window, slot_stride=size, nslots=nslots, timeout_ms=300)
    log(size, iters, window, r.status, r.done, round(r.gbps, 1), {k: v for k, v in qp.counters().items() if k in ("n_wqe", "n_cqe", "cursor", "retire_head", "sq_cons", "resv_head")})
    if not r.ok:
        log("engine", ctx.engine_stats())
        break
ctx.engine_stop(); log("stopped")
log("verify", ops.compare(src, dst))
