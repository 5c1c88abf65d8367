"""GPU tier-1 tests: the software HCA end to end (host-posted and GPU-posted)."""
import pytest
import torch

import rocnrdma_b200 as rn
from rocnrdma_b200 import ops, wire as W

pytestmark = pytest.mark.gpu


def _bufs(n, dev="cuda:0"):
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    dst = torch.zeros(n, dtype=torch.uint8, device=dev)
    ops.fill_random(src, seed=1234)
    return src, dst


def test_host_posted_write_small(ctx):
    src, dst = _bufs(4096)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=64, mem=W.MEM_HOST_PINNED)
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    qp.post_write(ms, md, 4096)
    wc = qp.scq.wait(1)[0]
    ctx.engine_stop()
    assert not wc.is_error and wc.opcode == W.CQE_REQ and wc.byte_cnt == 4096 and wc.qpn == qp.qpn
    assert torch.equal(src, dst)


@pytest.mark.parametrize("nbytes", [1, 7, 64, 1000, 4096, 65536 + 16, (1 << 20) + 3, 8 << 20])
def test_gpu_posted_write_sizes(ctx, nbytes):
    src, dst = _bufs(nbytes + 64)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=64)
    ctx.engine_start(ctas=16, idle_timeout_ms=3000)
    r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, nbytes, iters=3, window=2)
    ctx.engine_stop()
    assert r.ok, r.status
    assert r.done == [3]
    assert torch.equal(src[:nbytes], dst[:nbytes])
    assert int(dst[nbytes:].sum()) == 0, "engine wrote past the message"
    c = qp.counters()
    assert c["n_db_order_violations"] == 0 and c["n_err"] == 0 and c["n_wqe"] == 3


def test_gpu_posted_read_and_unaligned(ctx):
    n = 100_003
    src, dst = _bufs(n + 32)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=16)
    ctx.engine_start(ctas=8, idle_timeout_ms=3000)
    # READ: local = dst (+3 misalignment), remote = src (+5)
    lib_mr_l = rn.api.MemoryRegion(ctx, md.addr + 3, n, md.key, md.access)
    lib_mr_r = rn.api.MemoryRegion(ctx, ms.addr + 5, n, ms.key, ms.access)
    r = ops.rdma_stream(qp, W.OP_RDMA_READ, lib_mr_l, lib_mr_r, n - 8, iters=1)
    ctx.engine_stop()
    assert r.ok, r.status
    assert torch.equal(dst[3:3 + n - 8], src[5:5 + n - 8])


def test_bad_rkey_gives_error_cqe_and_flush(ctx):
    src, dst = _bufs(4096)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst, access=W.ACC_LOCAL_WRITE)  # no REMOTE_WRITE
    qp = ctx.loopback_qp(depth=16, mem=W.MEM_HOST_PINNED)
    ctx.engine_start(ctas=2, idle_timeout_ms=3000)
    qp.post_write(ms, md, 4096)
    qp.post_write(ms, md, 4096)
    wcs = qp.scq.wait(2)
    ctx.engine_stop()
    assert wcs[0].is_error and wcs[0].status == "REMOTE_ACCESS_ERR"
    assert wcs[1].is_error and wcs[1].status == "WR_FLUSH_ERR"
    assert qp.state == "ERR"
    assert int(dst.sum()) == 0


def test_out_of_bounds_and_revoked_mr(ctx):
    src, dst = _bufs(8192)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=16, mem=W.MEM_HOST_PINNED)
    ctx.engine_start(ctas=2, idle_timeout_ms=3000)
    qp.post_write(ms, md, 4096, dst_off=8192 - 100)     # runs off the end of the remote MR
    wc = qp.scq.wait(1)[0]
    assert wc.is_error and wc.status == "REMOTE_ACCESS_ERR"
    # revocation: a fresh QP, MR revoked under it (the cudaFree-while-registered case)
    qp2 = ctx.loopback_qp(depth=16, mem=W.MEM_HOST_PINNED)
    md.revoke()
    assert md.state == "REVOKED"
    qp2.post_write(ms, md, 64)
    wc = qp2.scq.wait(1)[0]
    ctx.engine_stop()
    assert wc.is_error and wc.status == "REMOTE_ACCESS_ERR"
    md.dereg()
    assert md.state == "FREE"


def test_send_recv_with_rnr(ctx):
    n = 32768
    src, dst = _bufs(n)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    cq_a, cq_b = ctx.create_cq(64, W.MEM_HOST_PINNED), ctx.create_cq(64, W.MEM_HOST_PINNED)
    qa = ctx.create_qp(cq_a, cq_a, 16, 16, W.MEM_HOST_PINNED)
    qb = ctx.create_qp(cq_b, cq_b, 16, 16, W.MEM_HOST_PINNED)
    qa.connect(qb)
    ctx.engine_start(ctas=4, idle_timeout_ms=3000, rnr_timeout_ms=2000)
    qa.post_send(ms, n, imm=0xabcd1234)          # receiver not ready yet -> engine retries
    import time
    time.sleep(0.05)
    qb.post_recv(md, n)
    wc_s = cq_a.wait(1)[0]
    wc_r = cq_b.wait(1)[0]
    ctx.engine_stop()
    assert not wc_s.is_error and not wc_r.is_error
    assert wc_r.opcode == W.CQE_RESP_SEND_IMM and wc_r.byte_cnt == n and wc_r.imm == 0xabcd1234
    assert torch.equal(src, dst)
    assert qa.counters()["n_rnr"] >= 1


def test_unsignaled_then_signaled(ctx):
    src, dst = _bufs(1 << 16)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=64)
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, 1024, iters=40, window=0, signal_every=8, slot_stride=1024,
                        nslots=40)
    ctx.engine_stop()
    assert r.ok and r.done == [40]
    c = qp.counters()
    assert c["n_wqe"] == 40 and c["n_cqe"] == 5
    assert torch.equal(src[:40 * 1024], dst[:40 * 1024])


def test_sq_wraparound_many_messages(ctx):
    src, dst = _bufs(1 << 16)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=8, cq_depth=16)
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, 256, iters=1000, window=4, slot_stride=256, nslots=256)
    ctx.engine_stop()
    assert r.ok and r.done == [1000], (r.status, r.done)
    assert torch.equal(src, dst)


def test_qp_state_machine(ctx):
    cq = ctx.create_cq(16)
    qp = ctx.create_qp(cq)
    assert qp.state == "RESET"
    with pytest.raises(rn._native.NativeError):
        qp.modify(W.QPS_RTS)
    qp.modify(W.QPS_INIT)
    with pytest.raises(rn._native.NativeError):
        qp.modify(W.QPS_RTR)          # not connected yet
    qp.connect_remote(qp.describe())
    qp.modify(W.QPS_RTR)
    qp.modify(W.QPS_RTS)
    assert qp.state == "RTS"
    qp.modify(W.QPS_RESET)
    assert qp.state == "RESET"


def test_engine_idle_watchdog(ctx):
    import time
    ctx.engine_start(ctas=2, idle_timeout_ms=200)
    assert ctx.engine_running
    time.sleep(0.6)
    assert not ctx.engine_running
    assert ctx.engine_stats()["exited_idle"] == 1


def test_many_ctas_one_sq_stress(ctx):
    """WQE-slot reservation / doorbell ordering under contention (128 CTAs, one SQ of 64)."""
    ctas, per = 128, 48
    n = ctas * per * 64
    src, dst = _bufs(n)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=64, cq_depth=128)
    torch.cuda.synchronize()
    ctx.engine_start(ctas=16, idle_timeout_ms=3000)
    try:
        r = ops.shared_post_stress(qp, ms, md, ctas=ctas, per_cta=per)
    finally:
        ctx.engine_stop()
    assert r["status"] == "OK" and r["posted"] == ctas * per, r
    assert torch.equal(src, dst)
    c = qp.counters()
    assert c["n_wqe"] == ctas * per + 1 and c["n_err"] == 0 and c["n_db_order_violations"] == 0, c
    assert c["resv_head"] == c["ready_head"] == c["sq_cons"] == c["retire_head"]


def test_rnr_retry_exceeded_when_no_receive_is_posted(ctx):
    src, _ = _bufs(4096)
    ms = ctx.reg_mr(src)
    cq_a, cq_b = ctx.create_cq(64, W.MEM_HOST_PINNED), ctx.create_cq(64, W.MEM_HOST_PINNED)
    qa = ctx.create_qp(cq_a, cq_a, 16, 16, W.MEM_HOST_PINNED)
    qb = ctx.create_qp(cq_b, cq_b, 16, 16, W.MEM_HOST_PINNED)
    qa.connect(qb)
    ctx.engine_start(ctas=2, idle_timeout_ms=3000, rnr_timeout_ms=50)
    qa.post_send(ms, 64)
    wc = cq_a.wait(1)[0]
    ctx.engine_stop()
    assert wc.is_error and wc.status == "RNR_RETRY_EXC_ERR"
    assert qa.state == "ERR" and qa.counters()["n_rnr"] >= 1


def test_gpu_posted_send_lands_in_posted_receive_buffers(ctx):
    n, k = 4096, 8
    src, dst = _bufs(n * k)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    cq_a, cq_b = ctx.create_cq(64), ctx.create_cq(64, W.MEM_HOST_PINNED)
    qa = ctx.create_qp(cq_a, cq_a, 16, 16)
    qb = ctx.create_qp(cq_b, cq_b, 16, 16, W.MEM_HOST_PINNED)
    qa.connect(qb)
    for i in range(k):
        qb.post_recv(md, n, off=i * n)
    torch.cuda.synchronize()
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    try:
        r = ops.rdma_stream(qa, W.OP_SEND, ms, None, n, iters=k, window=4, slot_stride=n, nslots=k)
        wcs = cq_b.wait(k)
    finally:
        ctx.engine_stop()
    assert r.ok and r.done == [k]
    assert all(w.opcode == W.CQE_RESP_SEND and w.byte_cnt == n for w in wcs)
    assert [w.wqe_counter for w in wcs] == list(range(k))          # receives are consumed in order
    assert torch.equal(src, dst)


def test_device_preposted_receives_and_gpu_posted_sends(ctx):
    """Both ends on the device: a consumer kernel posts its own receive WQEs (dev::post_recv) and polls the
    receive CQ while a poster kernel SENDs; no host-posted work request anywhere."""
    n, k = 8192, 6
    src, dst = _bufs(n * k)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    cq_a, cq_b = ctx.create_cq(64), ctx.create_cq(64)
    qa = ctx.create_qp(cq_a, cq_a, 16, 16)
    qb = ctx.create_qp(cq_b, cq_b, 16, 16)
    qa.connect(qb)
    torch.cuda.synchronize()
    ctx.engine_start(ctas=4, idle_timeout_ms=3000, rnr_timeout_ms=2000)
    try:
        view, rs = ops.recv_consume(qb, k, timeout_ms=3000, sync=False, stream=ctx.aux_stream, prepost_mr=md, prepost_bytes=n)
        r = ops.rdma_stream(qa, W.OP_SEND, ms, None, n, iters=k, window=2, slot_stride=n, nslots=k)
        rs.synchronize()
    finally:
        ctx.engine_stop()
    cons = ops.parse_recv(view)
    assert r.ok and cons["status"] == "OK" and cons["seen"] == k and cons["bytes"] == n * k, (r, cons)
    assert torch.equal(src, dst)


def test_batch_claim_error_in_the_middle_flushes_the_rest(ctx):
    """16 WQEs are visible when the engine starts, so one CAS claims them as a batch and a warp parses them:
    WQE 5 names a region without REMOTE_WRITE -> 0..4 complete, 5 fails with the IB syndrome, 6..15 are flushed
    in order, nothing of 5..15 moves."""
    n = 4096
    src, dst = _bufs(16 * n)
    bad = torch.zeros(n, dtype=torch.uint8, device="cuda:0")
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    mbad = ctx.reg_mr(bad, access=W.ACC_LOCAL_WRITE)
    qp = ctx.loopback_qp(depth=32)
    for i in range(16):
        if i == 5:
            qp.post_write(ms, mbad, n, src_off=i * n)
        else:
            qp.post_write(ms, md, n, src_off=i * n, dst_off=i * n)
    torch.cuda.synchronize()
    ctx.engine_run_oneshot(ctas=8)
    wcs = qp.scq.wait(16)
    assert [w.wqe_counter for w in wcs] == list(range(16))
    assert not any(w.is_error for w in wcs[:5])
    assert wcs[5].is_error and wcs[5].status == "REMOTE_ACCESS_ERR"
    assert all(w.is_error and w.status == "WR_FLUSH_ERR" for w in wcs[6:])
    assert qp.state == "ERR"
    assert torch.equal(dst[:5 * n], src[:5 * n]) and int(dst[5 * n:].sum()) == 0 and int(bad.sum()) == 0
    c = qp.counters()
    assert c["n_wqe"] == 16 and c["n_err"] == 11 and c["cursor"] == 16 and c["retire_head"] == 16


def test_burst_sends_match_receives_in_order(ctx):
    """Burst-posted SENDs go through the batch path; receive matching stays strictly in order."""
    n, k = 2048, 64
    src, dst = _bufs(n * k)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    cq_a, cq_b = ctx.create_cq(256), ctx.create_cq(256, W.MEM_HOST_PINNED)
    qa = ctx.create_qp(cq_a, cq_a, 128, 16)
    qb = ctx.create_qp(cq_b, cq_b, 16, 128, W.MEM_HOST_PINNED)
    qa.connect(qb)
    for i in range(k):
        qb.post_recv(md, n, off=i * n)
    torch.cuda.synchronize()
    ctx.engine_start(ctas=8, idle_timeout_ms=3000)
    try:
        r = ops.rdma_stream(qa, W.OP_SEND, ms, None, n, iters=k, window=64, burst=16, signal_every=16, slot_stride=n, nslots=k)
        wcs = cq_b.wait(k)
    finally:
        ctx.engine_stop()
    assert r.ok and r.done == [k]
    assert [w.wqe_counter for w in wcs] == list(range(k))
    assert all(w.opcode == W.CQE_RESP_SEND and w.byte_cnt == n for w in wcs)
    assert torch.equal(src, dst)
    assert qa.counters()["n_cqe"] == k // 16          # cq moderation: one send CQE per 16 WQEs


@pytest.mark.parametrize("nbytes", [64, 4096, 262144])
def test_burst_stream_many_qps_soak(ctx, nbytes):
    """8 QPs x 2048 burst-posted writes each: batch claims, warp retirement and the ticket spread all busy at once."""
    nq, iters = 8, 2048
    nslots = min(256, (32 << 20) // nbytes)
    per = nslots * nbytes
    src, dst = _bufs(nq * per)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qps = [ctx.loopback_qp(depth=256) for _ in range(nq)]
    torch.cuda.synchronize()
    ctx.engine_start(ctas=48, idle_timeout_ms=3000)
    try:
        r = ops.rdma_stream(qps, W.OP_RDMA_WRITE, ms, md, nbytes, iters=iters, window=128, burst=16, signal_every=16,
                            slot_stride=nbytes, nslots=nslots, stride=per, timeout_ms=5000)
    finally:
        ctx.engine_stop()
    assert r.ok and r.done == [iters] * nq, r
    assert ops.compare(src, dst) == 0
    for q in qps:
        c = q.counters()
        assert c["n_wqe"] == iters and c["n_err"] == 0 and c["n_db_order_violations"] == 0
        assert c["n_cqe"] == iters // 16 and c["n_bytes"] == iters * nbytes


def test_initialisation_queued_on_the_default_stream_precedes_the_transfer(ctx):
    """Buffers are usually initialised by torch on the legacy default stream, asynchronously.  Work still queued there when
    the engine becomes resident would be held behind the persistent kernel and run AFTER the transfer it was meant to
    precede (seen under compute-sanitizer, whose launches are slow: zero-fills landing on top of delivered data).
    engine_start() drains the legacy stream first; a long sleep kernel in front of the initialisation makes the
    late-initialisation case deterministic."""
    n = 4096
    src = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    dst = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=16)
    torch.cuda.synchronize()
    torch.cuda._sleep(200_000_000)              # ~0.1 s of default-stream work ahead of the initialisation
    ops.fill_random(src, seed=99)               # default stream, behind the sleep
    dst.zero_()
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    r = ops.rdma_stream(qp, W.OP_RDMA_WRITE, ms, md, n, iters=1)
    ctx.engine_stop()
    assert r.ok, r.status
    assert int(src.sum()) != 0 and torch.equal(src, dst)


def test_ops_without_engine_are_ordered_after_the_default_stream(ctx):
    """ops.* run on the context's non-blocking stream; with no engine resident they first wait for the caller's pending
    default-stream work, so tensors torch is still initialising are not raced."""
    M, N, K = 256, 256, 128
    torch.cuda.synchronize()
    torch.cuda._sleep(200_000_000)
    a = torch.randn(M, K, device="cuda:0").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda:0").to(torch.bfloat16)
    c = torch.zeros(M, N, device="cuda:0", dtype=torch.bfloat16)   # would land on top of the product if the GEMM ran first
    r = ops.gemm_send(ctx, a, b, c)
    assert r.ok
    assert torch.allclose(c.float(), a.float() @ b.float().T, rtol=2e-2, atol=0.5)


# ---------------------------------------------------------------- regression tests for the round-1 review findings
def test_counters_report_the_qp_state(ctx):
    """rn_qp_query used to leave `state` unwritten (the assignment sat behind a // comment): counters() said RESET for every QP."""
    src, dst = _bufs(4096)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=16, mem=W.MEM_HOST_PINNED)
    assert qp.state == "RTS" and qp.counters()["state"] == "RTS"
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    try:
        qp.post_raw(W.OP_RDMA_WRITE, ms.addr, ms.lkey, md.addr, 0xdead, 64)           # bad rkey -> error CQE -> QP in ERR
        wc = qp.scq.wait(1)[0]
    finally:
        ctx.engine_stop()
    assert wc.is_error
    assert qp.state == "ERR" and qp.counters()["state"] == "ERR"


def test_host_post_beyond_the_queue_depth_is_refused(ctx):
    """No flow control used to mean: the (depth + 1)-th post overwrote an unexecuted WQE.  Now it is ENOMEM, as ibv_post_send,
    and polling completions frees the slots again."""
    src, dst = _bufs(1 << 16)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=8, mem=W.MEM_HOST_PINNED)
    for i in range(8):                                                                  # no engine yet: nothing completes
        qp.post_write(ms, md, 4096, src_off=i * 4096, dst_off=i * 4096)
    with pytest.raises(rn._native.NativeError) as e:
        qp.post_write(ms, md, 4096)
    assert "-12" in str(e.value) or "ENOMEM" in str(e.value)
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    try:
        assert len(qp.scq.wait(8)) == 8
        for i in range(8):                                                              # a full window fits again
            qp.post_write(ms, md, 4096, src_off=(8 + i) * 4096, dst_off=(8 + i) * 4096)
        assert len(qp.scq.wait(8)) == 8
    finally:
        ctx.engine_stop()
    assert torch.equal(src, dst) and qp.counters()["n_err"] == 0
    rq = ctx.loopback_qp(depth=4, mem=W.MEM_HOST_PINNED)
    for _ in range(4):
        rq.post_recv(md, 64)
    with pytest.raises(rn._native.NativeError):
        rq.post_recv(md, 64)                                                            # the receive queue has a depth too


def test_unpolled_cq_overruns_instead_of_wrapping(ctx):
    """The engine used to reserve CQ slots without looking at the consumer index: an unpolled CQ wrapped over valid CQEs.
    Now the completion that would not fit raises a CQ overrun and the QP goes to ERR."""
    src, dst = _bufs(4096)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    cq = ctx.create_cq(4, W.MEM_HOST_PINNED)
    qp = ctx.create_qp(cq, cq, 32, 4, W.MEM_HOST_PINNED).connect()
    ctx.engine_start(ctas=4, idle_timeout_ms=3000)
    try:
        for _ in range(12):                                                             # 12 signalled writes, nobody polls a 4-entry CQ
            qp.post_write(ms, md, 64)
        import time
        t0 = time.time()
        while qp.state != "ERR" and time.time() - t0 < 3:
            time.sleep(0.01)
    finally:
        ctx.engine_stop()
    assert qp.state == "ERR"
    good = [w for w in cq.poll(16) if not w.is_error]
    assert len(good) <= 4                                                               # what is there is intact: never more than the ring holds


def test_shared_send_and_receive_cq_with_write_imm(ctx):
    """cq_poll_once used to take every CQE for a send completion of its QP.  With one CQ for both directions a receive
    completion (WRITE_IMM) was swallowed and its wqe_counter was applied to sq_cons.  The device poster must only consume
    its own requester CQEs; the receive side gets its arrivals."""
    chunk = 1 << 16
    x = torch.randn(1 << 19, device="cuda:0").to(torch.bfloat16)
    nb = ops.staging_bytes(x.numel(), chunk)
    stg = torch.zeros(nb, dtype=torch.uint8, device="cuda:0")
    rem = torch.zeros(nb, dtype=torch.uint8, device="cuda:0")
    y = torch.zeros_like(x)
    smr, rmr = ctx.reg_mr(stg), ctx.reg_mr(rem)
    qp = ctx.loopback_qp(depth=64, shared_cq=True)
    n_rec = x.numel() // chunk
    for _ in range(n_rec):
        qp.post_recv(rmr, 0)
    _, s_rx = ctx.streams(2)
    torch.cuda.synchronize()
    ctx.engine_start(ctas=8, idle_timeout_ms=3000)
    try:
        # receiver first (it waits on the shared CQ from the device), then the poster: each must leave the other's CQEs alone
        view, _ = ops.unpack_fp8(ctx, rem, y, chunk_elems=chunk, qp=qp, timeout_ms=4000, sync=False, stream=s_rx)
        pr = ops.pack_fp8_write(ctx, x, smr, qp=qp, dst_mr=rmr, chunk_elems=chunk, with_imm=True, signal_every=1, timeout_ms=4000)
        s_rx.synchronize()
    finally:
        ctx.engine_stop()
    assert pr.ok, pr.status
    assert torch.equal(rem, ops.ref_pack_fp8(x, chunk))
    assert torch.equal(y, ops.ref_unpack_fp8(rem, x.numel(), chunk))
    c = qp.counters()
    assert c["n_err"] == 0 and c["sq_cons"] <= c["resv_head"]


def test_reset_forgets_stale_ready_flags_and_wqes(ctx):
    """RESET rewound the indices but kept ready_flags[] and the old WQE bytes: after reuse the shared submit rang the doorbell
    over unwritten slots and the engine re-executed stale descriptors.  Post a few through the shared path, reset, reconnect,
    post fewer: exactly the new ones execute."""
    src, dst = _bufs(1 << 16)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    qp = ctx.loopback_qp(depth=64)
    torch.cuda.synchronize()
    ctx.engine_start(ctas=8, idle_timeout_ms=3000)
    try:
        r = ops.shared_post_stress(qp, ms, md, ctas=4, per_cta=6)                       # 24 WQEs through sq_submit_shared
        assert r["status"] == "OK" and r["posted"] == 24
    finally:
        ctx.engine_stop()
    qp.modify(W.QPS_RESET)
    qp.connect()
    assert qp.state == "RTS"
    dst.zero_()
    torch.cuda.synchronize()
    ctx.engine_start(ctas=8, idle_timeout_ms=3000)
    try:
        r = ops.shared_post_stress(qp, ms, md, ctas=2, per_cta=3)                       # 6 WQEs: slots 6..23 still hold the old bytes
        assert r["status"] == "OK" and r["posted"] == 6
    finally:
        ctx.engine_stop()
    c = qp.counters()
    # 6 writes + the kernel's flush NOP; a stale slot executed again would show as more WQEs (and as bytes in dst)
    assert c["n_wqe"] == 7 and c["n_err"] == 0 and c["retire_head"] == 7, c
