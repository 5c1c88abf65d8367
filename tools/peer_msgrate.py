"""Message rate and WQE lifecycle of a QP whose responder lives on ANOTHER GPU (config 4 wiring):
plain writes vs writes with immediate (receive matching + receive CQE on the peer), window 1 and 32."""
import json, sys, os, statistics as st
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rocnrdma_b200 import ops, wire as W
from rocnrdma_b200.api import Context

tx, rx = Context(0), Context(1)
tx.enable_peer(1)
N = 8 << 20
src = torch.empty(N, dtype=torch.uint8, device="cuda:0"); dst = torch.empty(N, dtype=torch.uint8, device="cuda:1")
ms, md = tx.reg_mr(src), rx.reg_mr(dst)
cq_a, cq_b = tx.create_cq(1024), rx.create_cq(4096)
qa = tx.create_qp(cq_a, cq_a, 256, 16); qb = rx.create_qp(cq_b, cq_b, 16, 2048)
qa.connect(qb); qa.set_flags(sys_scope=True, trace=True)
ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tx.engine_start(ctas=ctas, idle_timeout_ms=3000)
out = {}
for op, name in [(W.OP_RDMA_WRITE, "write"), (W.OP_RDMA_WRITE_IMM, "write_imm")]:
    for size in [0, 64, 65536]:
        if size == 0 and op == W.OP_RDMA_WRITE: continue
        for window in [1, 32]:
            iters = 256
            if op == W.OP_RDMA_WRITE_IMM:
                for _ in range(iters): qb.post_recv(md, 0)
            r = ops.rdma_stream(qa, op, ms, md, size, iters=iters, window=window, burst=max(1, window // 4), signal_every=max(1, window // 4),
                                slot_stride=max(size, 64), nslots=32, timeout_ms=3000)
            tr = [t for t in qa.read_trace(256) if t["post"] and t["seen"] >= t["post"]]
            seg = lambda a, b: round(st.median([(t[b] - t[a]) / 1e3 for t in tr]), 2) if tr else None
            row = dict(status=r.status, us_per_msg=round(r.us_per_msg, 2), post_to_claim=seg("post", "claim"), claim_to_parsed=seg("claim", "parsed"),
                       parsed_to_copied=seg("parsed", "copied"), copied_to_cqe=seg("copied", "cqe"), cqe_to_seen=seg("cqe", "seen"), total=seg("post", "seen"))
            out[f"{name}_{size}B_w{window}"] = row
            print(name, size, window, row, flush=True)
            while rx_polled := cq_b.poll(256): pass
tx.engine_stop()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(ctas=ctas, rows=out), open("gpurun_out/peer_msgrate.json", "w"), indent=1)
