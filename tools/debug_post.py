import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
T0 = time.time()
def log(*a): print(f"[{time.time()-T0:7.2f}s]", *a, flush=True)
import torch
log("torch imported")
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W
ctx = rn.Context(0); log("ctx")
src = torch.empty(1 << 20, dtype=torch.uint8, device="cuda"); dst = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
ops.fill_random(src, 3); torch.cuda.synchronize(); log("bufs")
ms, md = ctx.reg_mr(src), ctx.reg_mr(dst); log("mrs")
qp = ctx.loopback_qp(depth=64); log("qp", qp.counters())
side = torch.cuda.Stream()
for mode in ["legacy", "side"]:
    ctx.engine_start(ctas=4, idle_timeout_ms=1500); log(mode, "engine started", ctx.engine_running)
    time.sleep(0.2)
    log("stats before", ctx.engine_stats())
    st = None if mode == "legacy" else side
    out, keep = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, 4096, iters=2, window=0, timeout_ms=500, stream=st, sync=False); st = keep
    log("poster launched")
    time.sleep(0.3)
    log("stats during", ctx.engine_stats(), "running", ctx.engine_running)
    (st or torch.cuda.current_stream()).synchronize()
    log("poster done", ops.rdma.parse_stream_out(out, 1, 4096))
    log("stats after", ctx.engine_stats(), qp.counters())
    ctx.engine_stop(); log("engine stopped")
log("equal", torch.equal(src[:4096], dst[:4096]))
