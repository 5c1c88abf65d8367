"""WQE lifecycle breakdown (device %globaltimer stamps) for small GPU-posted writes."""
import json, sys, os, statistics as st
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W
ctx = rn.Context(0)
N = 64 << 20
src = torch.empty(N, dtype=torch.uint8, device="cuda"); dst = torch.empty(N, dtype=torch.uint8, device="cuda")
ops.fill_random(src, 7); torch.cuda.synchronize()
ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
qp = ctx.loopback_qp(depth=256)
qp.set_flags(trace=True)
ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx.engine_start(ctas=ctas, idle_timeout_ms=2000)
out = {}
for size in [64, 4096, 65536, 1 << 20]:
    for window in [1, 16]:
        ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=64, window=window, slot_stride=size, nslots=32)
        r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, size, iters=192, window=window, slot_stride=size, nslots=32)
        tr = qp.read_trace(256)
        tr = [t for t in tr if t["post"] and t["seen"] >= t["post"]]
        seg = lambda a, b: round(st.median([(t[b] - t[a]) / 1e3 for t in tr]), 2)
        row = dict(us_per_msg=round(r.us_per_msg, 2), gbps=round(r.gbps, 2), post_to_claim=seg("post", "claim"),
                   claim_to_parsed=seg("claim", "parsed"), parsed_to_copied=seg("parsed", "copied"),
                   copied_to_cqe=seg("copied", "cqe"), cqe_to_seen=seg("cqe", "seen"), total=seg("post", "seen"))
        out[f"{size}B_w{window}"] = row
        print(size, window, row, flush=True)
ctx.engine_stop()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(ctas=ctas, rows=out), open("gpurun_out/latency_trace.json", "w"), indent=1)
print(qp.counters())
