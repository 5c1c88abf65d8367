"""2-GPU paths (NVLink wire): one-sided write into a peer GPU, and config 4 (GEMM on GPU0 -> panels on GPU1)."""
import pytest
import torch

import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _need2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_write_into_peer_gpu_over_nvlink():
    _need2()
    ctx = rn.Context(0)
    n = 32 << 20
    src = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    dst = torch.zeros(n, dtype=torch.uint8, device="cuda:1")
    ops.fill_random(src, 99)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)          # peer HBM: registration enables the NVLink mapping
    qp = ctx.loopback_qp(depth=32)
    qp.set_flags(sys_scope=True)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    ctx.engine_start(ctas=32, idle_timeout_ms=3000)
    try:
        r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, n, iters=4, window=2)
        rd = ops.rdma_stream(qp, W.OP_RDMA_READ, ms, md, 1 << 20, iters=1)
    finally:
        ctx.engine_stop()
    assert r.ok and rd.ok
    assert torch.equal(src.cpu(), dst.cpu())
    assert r.gbps > 100, f"NVLink write only {r.gbps:.1f} GB/s"
    ctx.close()


def test_config4_gemm_on_gpu0_panels_land_on_gpu1():
    _need2()
    from rocnrdma_b200.models import sendrecv_gemm
    res = sendrecv_gemm.run(M=1024, N=1024, K=1024, engine_ctas=16, reps=2)
    assert res.ok, res
    assert res.verified, "GPU1's buffer differs from what GPU0 computed"
    assert res.consumer["seen"] == 8 and res.consumer["bytes"] == 2 * 1024 * 1024


def test_cross_process_ring_over_cuda_ipc(tmp_path):
    """One process per GPU, QPs connected through parallel.connect_ring; every rank writes into the next one."""
    _need2()
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "ring.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "bench", "ring.py"), "--sizes", "1m,64m", "--engine-ctas", "16", "--out", str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=root)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(out.read_text())
    assert res["verified"] and res["world"] == 2
    assert res["rows"][-1]["per_gpu_gbps"] > 200


def test_config4_direct_mode_epilogue_stores_over_nvlink():
    """GEMM epilogue writes straight into GPU1's buffer; only the per-panel signal uses the queue pair."""
    _need2()
    from rocnrdma_b200.models import sendrecv_gemm
    res = sendrecv_gemm.run(M=1024, N=1024, K=1024, reps=2, mode="direct")
    assert res.ok, res
    assert res.verified
    assert res.consumer["seen"] == 8 and res.consumer["bytes"] == 0      # zero-length signals


def test_chained_gemms_with_the_wire_fused_into_both():
    """GPU0's GEMM sends fp8 panel records panel by panel; GPU1's block-scaled GEMM multiplies them where they land and
    starts each tile when its panel has arrived.  Both schedules (fused, one-after-the-other) must give the same,
    verified, result."""
    _need2()
    from rocnrdma_b200.models import sendrecv_gemm as SG
    r = SG.run_chain(2048, 1024, 512, 1024, reps=1)
    assert r["verified"] and r["fused_us"] and r["sequential_us"], r


@pytest.mark.parametrize("mode", ["engine", "direct"])
def test_pack_on_gpu0_unpack_on_gpu1(mode):
    """K3 -> NVLink -> K5 with no host step between: the peer's unpack kernel waits on its receive CQ from the device.
    engine: records staged locally, moved by the DMA engine; direct: the pack kernel stores into the peer's registered
    buffer itself and only announces each record through the queue pair."""
    _need2()
    from rocnrdma_b200.models import sendrecv_pack as SP
    r = SP.run(1 << 24, 1 << 20, mode=mode, reps=1)
    assert r["verified"] and r["fused_us"] and r["sequential_us"], r
