"""The ConnectX backend (csrc/verbs/verbs_dl.cc) end to end against the in-tree mock rdma-core provider
(csrc/mockverbs): everything here runs on a CPU-only box.  Covers BASELINE config 1 (host-DRAM ibv_reg_mr +
RDMA write loopback between two ports), the verbs error model, the three registration routes of GPU memory
(peer-memory client present / absent / the in-tree b200p2p bridge, dma-buf), and the mlx5dv raw-queue contract
a GPU poster relies on -- driven here by a host stand-in that writes WQE bytes, doorbell record and doorbell
register by hand, directly and through the CPU doorbell proxy."""
import ctypes as C
import mmap
import os
import time

import numpy as np
import pytest

import rocnrdma_b200 as rn
from rocnrdma_b200 import _native as N, wire as W

H = W.MEM_HOST_PINNED


@pytest.fixture(scope="module")
def lib():
    l = N.load()
    assert l.rn_verbs_compiled() == 1
    assert l.rn_verbs_available() >= 2, l.rn_verbs_why()
    assert l.rn_verbs_is_mock() == 1
    return l


@pytest.fixture
def pair(lib):
    """Two host-only contexts on two mock HCAs with a connected host-posted QP each (NIC0 <-> NIC1)."""
    c0, c1 = rn.Context(device=None, wire="verbs", nic=0), rn.Context(device=None, wire="verbs", nic=1)
    q0 = c0.create_qp(c0.create_cq(64, H), c0.create_cq(64, H), 16, 16, H)
    q1 = c1.create_qp(c1.create_cq(64, H), c1.create_cq(64, H), 16, 16, H)
    q0.connect(q1)
    yield c0, c1, q0, q1
    c0.close()
    c1.close()


def _bufs(n=1 << 16):
    rng = np.random.default_rng(7)
    return rng.integers(0, 255, n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)


def test_discovery_and_auto_never_picks_the_mock(lib):
    name = C.create_string_buffer(64)
    assert lib.rn_verbs_device_name(0, name, 64) == 0 and name.value == b"mock_mlx5_0"
    assert lib.rn_verbs_device_name(9, name, 64) != 0
    assert rn.api.resolve_wire("auto") == "softhca"
    assert rn.api.resolve_wire("verbs") == "verbs"
    with pytest.raises(ValueError):
        rn.api.resolve_wire("ethernet")
    with pytest.raises(ValueError):
        rn.Context(device=None, wire="softhca")


def test_host_dram_write_read_send_between_two_ports(pair):
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma, mb = c0.reg_mr(a), c1.reg_mr(b)
    assert c0.nic == "mock_mlx5_0" and c1.nic == "mock_mlx5_1" and ma.mode == "peermem"
    assert q0.state == "RTS" and q1.state == "RTS"
    q0.post_write(ma, mb, 4096)
    wc = q0.scq.wait(1)[0]
    assert wc.status == "OK" and wc.opcode == W.CQE_REQ and wc.byte_cnt == 4096 and wc.qpn == q0.qpn
    assert (a[:4096] == b[:4096]).all() and not b[4096:].any()
    # SEND with immediate into a posted receive buffer; completions on both sides
    q1.post_recv(mb, 1024, off=8192)
    q0.post_send(ma, 512, src_off=100, imm=0xDEADBEEF)
    assert q0.scq.wait(1)[0].status == "OK"
    wr = q1.rcq.wait(1)[0]
    assert wr.opcode == W.CQE_RESP_SEND_IMM and wr.imm == 0xDEADBEEF and wr.byte_cnt == 512 and wr.qpn == q1.qpn
    assert (a[100:612] == b[8192:8192 + 512]).all()
    # RDMA READ pulls the other way
    q0.post_read(ma, mb, 256, local_off=32768, remote_off=0)
    assert q0.scq.wait(1)[0].status == "OK"
    assert (a[32768:32768 + 256] == b[:256]).all()
    # WRITE_WITH_IMM consumes a receive WQE but ignores its buffer
    q1.post_recv(mb, 16, off=0)
    q0.post_write(ma, mb, 64, dst_off=20000, imm=77)
    assert q0.scq.wait(1)[0].status == "OK"
    wr = q1.rcq.wait(1)[0]
    assert wr.opcode == W.CQE_RESP_WR_IMM and wr.imm == 77
    c = q0.counters()
    assert c["nic"] == "mock" and c["n_wqe"] == 4 and c["n_err"] == 0 and c["n_db_order_violations"] == 0


def test_unsignaled_writes_complete_with_the_next_signaled_one(pair):
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma, mb = c0.reg_mr(a), c1.reg_mr(b)
    for i in range(7):
        q0.post_write(ma, mb, 1000, src_off=i * 1000, dst_off=i * 1000, signaled=False)
    q0.post_write(ma, mb, 1000, src_off=7000, dst_off=7000)
    wcs = q0.scq.wait(1)
    assert len(wcs) == 1 and wcs[0].status == "OK"
    time.sleep(0.01)
    assert q0.scq.poll() == []
    assert (a[:8000] == b[:8000]).all()
    # the 8 slots are free again: a full window can be posted without ENOMEM
    for i in range(16):
        q0.post_write(ma, mb, 8, signaled=(i == 15))
    assert q0.scq.wait(1)[0].status == "OK"


def test_send_queue_overflow_is_refused(pair):
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma = c0.reg_mr(a)
    mb_bad = rn.api.MemoryRegion(c1, 0x1000, 64, 0x77777, 7)
    # nobody polls: 16 slots fill up (the first WQE fails and the rest flush, but the host has not reaped them)
    for _ in range(16):
        q0.post_write(ma, mb_bad, 8)
    with pytest.raises(N.NativeError, match="post_send"):
        q0.post_write(ma, mb_bad, 8)
    assert len(q0.scq.wait(16)) == 16


def test_bad_rkey_fails_the_qp_then_reset_and_reconnect(pair):
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma, mb = c0.reg_mr(a), c1.reg_mr(b, access=W.ACC_LOCAL_WRITE)       # no REMOTE_WRITE
    q0.post_write(ma, mb, 64)
    q0.post_write(ma, mb, 64)
    w = q0.scq.wait(2)
    assert [x.status for x in w] == ["REMOTE_ACCESS_ERR", "WR_FLUSH_ERR"] and all(x.is_error for x in w)
    assert q0.state == "ERR" and not b.any()
    # out of bounds on a good key is the same class of error
    q0.modify(W.QPS_RESET)
    q1.modify(W.QPS_RESET)
    assert q0.state == "RESET"
    q0.connect(q1)
    mb2 = c1.reg_mr(b)
    q0.post_write(ma, mb2, 128, dst_off=(1 << 16) - 64)
    assert q0.scq.wait(1)[0].status == "REMOTE_ACCESS_ERR"
    q0.modify(W.QPS_RESET); q1.modify(W.QPS_RESET); q0.connect(q1)
    q0.post_write(ma, mb2, 128)
    assert q0.scq.wait(1)[0].status == "OK" and (a[:128] == b[:128]).all()
    # a deregistered key stops translating
    mb2.dereg()
    q0.post_write(ma, mb2, 128)
    assert q0.scq.wait(1)[0].status == "REMOTE_ACCESS_ERR"


def test_receiver_not_ready_retries_then_gives_up(pair, lib):
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma, mb = c0.reg_mr(a), c1.reg_mr(b)
    lib.rn_verbs_mock_set_rnr_timeout_ms(2000)
    q0.post_send(ma, 256)
    time.sleep(0.05)
    assert q0.scq.poll() == []                   # held back: no receive buffer yet
    q1.post_recv(mb, 256)
    assert q0.scq.wait(1)[0].status == "OK" and q1.rcq.wait(1)[0].byte_cnt == 256
    assert q0.counters()["n_rnr"] >= 1 and (a[:256] == b[:256]).all()
    lib.rn_verbs_mock_set_rnr_timeout_ms(30)
    q0.post_send(ma, 256)
    assert q0.scq.wait(1)[0].status == "RNR_RETRY_EXC_ERR"
    lib.rn_verbs_mock_set_rnr_timeout_ms(500)
    # a receive buffer that is too small is the responder's complaint
    q0.modify(W.QPS_RESET); q1.modify(W.QPS_RESET); q0.connect(q1)
    q1.post_recv(mb, 16)
    q0.post_send(ma, 256)
    assert q0.scq.wait(1)[0].status == "REMOTE_INVAL_REQ_ERR"
    assert q1.rcq.wait(1)[0].is_error


def test_dmabuf_registration_with_a_host_exporter(pair):
    """ibv_reg_dmabuf_mr: a memfd stands in for the dma-buf (CPU-mappable exporter); iova is an arbitrary address
    the WQEs use, the NIC reaches the pages through the fd."""
    c0, c1, q0, q1 = pair
    lib = c0._lib
    fd = os.memfd_create("rn-dmabuf", 0)
    os.ftruncate(fd, 1 << 16)
    a, _ = _bufs()
    ma = c0.reg_mr(a)
    lk, rk = C.c_uint32(), C.c_uint32()
    iova = 0x7E00_0000_0000
    h = lib.rn_verbs_reg_mr(c1._vdev, iova, 1 << 15, 1, fd, 4096, 7, C.byref(lk), C.byref(rk))
    assert h, lib.rn_verbs_why()
    remote = rn.api.MemoryRegion(c1, iova, 1 << 15, lk.value, 7, rkey=rk.value)
    q0.post_write(ma, remote, 5000, dst_off=100)
    assert q0.scq.wait(1)[0].status == "OK"
    with mmap.mmap(fd, 1 << 16) as m:
        got = np.frombuffer(m, dtype=np.uint8, count=5000, offset=4096 + 100).copy()
    assert (got == a[:5000]).all()
    # too short an fd is refused
    assert not lib.rn_verbs_reg_mr(c1._vdev, iova, 1 << 20, 1, fd, 0, 7, C.byref(lk), C.byref(rk))
    assert b"ibv_reg_dmabuf_mr" in lib.rn_verbs_why()
    assert lib.rn_verbs_dereg_mr(h) == 0
    os.close(fd)


# ---------------------------------------------------------------- the raw-queue contract a GPU poster relies on
class HandPoster:
    """What hca/post.cuh does, from the host: WQE bytes into the mlx5dv send queue, big-endian doorbell record,
    8-byte store of the ctrl segment's head into the doorbell register; CQEs read straight from the ring."""

    def __init__(self, qp, doorbell_addr=None):
        self.lib, self.raw, self.qpn = qp.ctx._lib, qp.raw_queues(), qp.qpn
        self.head = 0
        self.ci = 0
        self.db = doorbell_addr or self.raw["bf_reg"]

    def post(self, opcode, laddr, lkey, raddr, rkey, nbytes, signaled=True, ring=True, record=True):
        r = self.raw
        assert r["sq_stride"] == 64 and r["cq_cqe_size"] == 64
        w = (C.c_uint8 * 64)()
        self.lib.rn_wire_build_wqe(w, opcode, self.head & 0xFFFF, self.qpn, laddr, lkey, raddr, rkey, nbytes,
                                   W.CTRL_CQ_UPDATE if signaled else 0, 0)
        C.memmove(r["sq_buf"] + (self.head % r["sq_wqe_cnt"]) * 64, w, 64)
        self.head += 1
        if record:
            C.c_uint32.from_address(r["dbrec"] + 4).value = int.from_bytes((self.head & 0xFFFF).to_bytes(4, "big"), "little")
        if ring:
            C.c_uint64.from_address(self.db).value = int.from_bytes(bytes(w[:8]), "little")

    def poll(self, timeout=2.0):
        r = self.raw
        t0 = time.time()
        slot = r["cq_buf"] + (self.ci % r["cq_cqe_cnt"]) * 64
        while True:
            op_own = C.c_uint8.from_address(slot + 63).value
            if (op_own >> 4) != W.CQE_INVALID and (op_own & 1) == ((self.ci // r["cq_cqe_cnt"]) & 1):
                break
            assert time.time() - t0 < timeout, "no CQE"
            time.sleep(0.0005)
        wc = N.RnWc()
        self.lib.rn_wire_decode_cqe((C.c_uint8 * 64).from_address(slot), C.byref(wc))
        self.ci += 1
        C.c_uint32.from_address(r["cq_dbrec"]).value = int.from_bytes((self.ci & 0xFFFFFF).to_bytes(4, "big"), "little")
        return wc


def test_raw_mlx5dv_queues_driven_by_a_foreign_poster(pair):
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma, mb = c0.reg_mr(a), c1.reg_mr(b)
    raw = q0.raw_queues()
    assert raw["sq_wqe_cnt"] == 16 and raw["rq_stride"] == 16 and raw["bf_size"] == 256 and raw["qpn"] == q0.qpn
    assert raw["cq_cqe_cnt"] == 128 and raw["rcq_buf"] != raw["cq_buf"]
    assert raw["sq_buf"] == raw["rq_buf"] + 16 * 16        # rdma-core lays one buffer out as [RQ | SQ]
    hp = HandPoster(q0)
    # 40 messages through a 16-deep ring: indices, wrap and the CQE owner bit all go round
    for i in range(40):
        hp.post(W.OP_RDMA_WRITE, ma.addr + i * 100, ma.lkey, mb.addr + i * 100, mb.rkey, 100)
        wc = hp.poll()
        assert not wc.is_error and wc.qpn == q0.qpn and wc.wqe_counter == (i & 0xFFFF) and wc.byte_cnt == 100
    assert (a[:4000] == b[:4000]).all()
    hp.post(W.OP_RDMA_READ, ma.addr + 50000, ma.lkey, mb.addr, mb.rkey, 64)
    assert hp.poll().wqe_opcode == W.OP_RDMA_READ and (a[50000:50064] == b[:64]).all()
    st = q0.counters()
    assert st["n_db_order_violations"] == 0 and st["n_doorbells"] == 41


def test_doorbell_that_outruns_its_record_is_counted(pair):
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma, mb = c0.reg_mr(a), c1.reg_mr(b)
    hp = HandPoster(q0)
    hp.post(W.OP_RDMA_WRITE, ma.addr, ma.lkey, mb.addr, mb.rkey, 64, record=False)    # register without the record: the bug
    time.sleep(0.05)
    assert q0.counters()["n_db_order_violations"] == 1 and not b.any()                  # and the NIC did not execute it
    hp.head = 0
    hp.post(W.OP_RDMA_WRITE, ma.addr, ma.lkey, mb.addr, mb.rkey, 64, ring=False)        # record, then a fresh doorbell value
    C.c_uint64.from_address(hp.raw["bf_reg"] + 256).value = 1                           # (the second BlueFlame register)
    assert not hp.poll().is_error and (a[:64] == b[:64]).all()


def test_cpu_doorbell_proxy_forwards_the_mailbox(pair):
    """The fallback when the UAR page cannot be mapped into the GPU: the poster stores the doorbell value in a
    mailbox, a host thread forwards it to the register."""
    c0, c1, q0, q1 = pair
    a, b = _bufs()
    ma, mb = c0.reg_mr(a), c1.reg_mr(b)
    mbox = C.c_uint64()
    assert c0._lib.rn_verbs_db_proxy_attach(q0._vq, C.byref(mbox)) == 0, c0._lib.rn_verbs_why()
    hp = HandPoster(q0, doorbell_addr=mbox.value)
    for i in range(5):
        hp.post(W.OP_RDMA_WRITE, ma.addr + i * 64, ma.lkey, mb.addr + i * 64, mb.rkey, 64)
        assert not hp.poll().is_error
    assert (a[:320] == b[:320]).all()
    assert q0.counters()["db_proxy_forwarded"] == 5
    with pytest.raises(N.NativeError, match="owned by the GPU poster"):
        q0.post_write(ma, mb, 8)                 # host posting on a QP whose queues were handed over is refused


# ---------------------------------------------------------------- GPU memory registration routes (no GPU needed)
def _fake_gpu_buffer(n=1 << 17):
    """64 KiB-aligned anonymous memory the mock is told to treat as GPU memory (ib_core cannot pin it)."""
    m = mmap.mmap(-1, n + (1 << 16))
    base = C.addressof(C.c_char.from_buffer(m))
    va = (base + 65535) & ~65535
    return m, va


def test_gpu_pointer_needs_a_peer_memory_client(pair, monkeypatch):
    c0, c1, q0, q1 = pair
    lib = c0._lib
    keep, va = _fake_gpu_buffer()
    assert lib.rn_verbs_mock_declare_gpu_range(va, 1 << 17) == 0
    monkeypatch.setenv("ROCNRDMA_MOCK_PEERMEM", "0")
    with pytest.raises(N.NativeError, match="no peer-memory client"):
        c1.reg_mr((va, 4096), mode="peermem")
    monkeypatch.setenv("ROCNRDMA_MOCK_PEERMEM", "1")
    mr = c1.reg_mr((va, 4096), mode="peermem")
    assert mr.state == "PINNED"
    mr.dereg()
    assert mr.state == "FREE"
    lib.rn_verbs_mock_gpu_free(va)
    del keep


def test_registration_through_the_b200p2p_bridge_and_driver_revocation(pair, monkeypatch):
    """ibv_reg_mr(gpu_va) the way the reference makes it work (amdp2p.c:112-264): ib_core asks the peer-memory
    client, here kmod/b200p2p.c itself (compiled against kmod/shim); the MR translates through the BUS addresses
    its dma_map returned; freeing the memory runs its free callback, which invalidates the MR (amdp2p.c:88-109)."""
    c0, c1, q0, q1 = pair
    lib = c0._lib
    monkeypatch.setenv("ROCNRDMA_MOCK_PEERMEM", "b200p2p")
    keep, va = _fake_gpu_buffer()
    assert lib.rn_verbs_mock_declare_gpu_range(va, 1 << 17) == 0
    assert lib.rn_verbs_mock_bridge_status() == b"ok", lib.rn_verbs_mock_bridge_status()
    a, _ = _bufs(1 << 17)
    ma = c0.reg_mr(a)
    # an unaligned sub-range: the module pins the enclosing 64 KiB pages itself
    off, n = 70_000, 50_000
    mr = c1.reg_mr((va + off, n), mode="peermem")
    q0.post_write(ma, mr, n)
    assert q0.scq.wait(1)[0].status == "OK"
    got = np.frombuffer(keep, dtype=np.uint8, count=n, offset=(va - C.addressof(C.c_char.from_buffer(keep))) + off)
    assert (got == a[:n]).all()
    # one byte past the registered range is not covered even though the pin is
    q0.post_write(ma, mr, 16, dst_off=n - 8)
    assert q0.scq.wait(1)[0].status == "REMOTE_ACCESS_ERR"
    q0.modify(W.QPS_RESET); q1.modify(W.QPS_RESET); q0.connect(q1)
    # "cudaFree" under the live MR: the driver revokes, the module invalidates, the NIC stops translating
    assert lib.rn_verbs_mock_gpu_free(va) >= 1
    q0.post_write(ma, mr, 64)
    assert q0.scq.wait(1)[0].status == "REMOTE_ACCESS_ERR"
    mr.dereg()                                    # dma_unmap / put_pages are no-ops after the revoke; release runs
    # and a host pointer never goes to the client at all
    _, b = _bufs()
    assert c1.reg_mr(b).state == "PINNED"
    del got
