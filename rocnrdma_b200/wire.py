"""Pure-Python mirror of ``csrc/wire/mlx5_wire.h`` (mlx5 WQE / CQE wire format).

Used by the tier-0 unit tests (no GPU, no NIC: SURVEY.md section 4.3) to check the native
encoders byte for byte, and by tools that want to pretty-print a queue dump.
All multi-byte fields are big-endian on the wire.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

WQEBB = 64
CQE_SIZE = 64

OP_NOP = 0x00
OP_SEND_INVAL = 0x01
OP_RDMA_WRITE = 0x08
OP_RDMA_WRITE_IMM = 0x09
OP_SEND = 0x0A
OP_SEND_IMM = 0x0B
OP_RDMA_READ = 0x10
OP_ATOMIC_CS = 0x11
OP_ATOMIC_FA = 0x12

CTRL_SOLICITED = 0x02
CTRL_CQ_UPDATE = 0x08
CTRL_FENCE = 0x80

CQE_REQ = 0x0
CQE_RESP_WR_IMM = 0x1
CQE_RESP_SEND = 0x2
CQE_RESP_SEND_IMM = 0x3
CQE_RESP_SEND_INV = 0x4
CQE_REQ_ERR = 0xD
CQE_RESP_ERR = 0xE
CQE_INVALID = 0xF

SYNDROMES = {
    0x00: "OK", 0x01: "LOCAL_LENGTH_ERR", 0x02: "LOCAL_QP_OP_ERR", 0x04: "LOCAL_PROT_ERR",
    0x05: "WR_FLUSH_ERR", 0x06: "MW_BIND_ERR", 0x10: "BAD_RESP_ERR", 0x11: "LOCAL_ACCESS_ERR",
    0x12: "REMOTE_INVAL_REQ_ERR", 0x13: "REMOTE_ACCESS_ERR", 0x14: "REMOTE_OP_ERR",
    0x15: "TRANSPORT_RETRY_EXC_ERR", 0x16: "RNR_RETRY_EXC_ERR", 0x22: "REMOTE_ABORTED_ERR",
}
SYN = {v: k for k, v in SYNDROMES.items()}

ACC_LOCAL_WRITE, ACC_REMOTE_WRITE, ACC_REMOTE_READ, ACC_REMOTE_ATOMIC = 1, 2, 4, 8
ACC_ALL = 15

MEM_DEVICE, MEM_HOST_PINNED, MEM_PEER = 0, 1, 2

QPS_RESET, QPS_INIT, QPS_RTR, QPS_RTS, QPS_SQD, QPS_SQE, QPS_ERR = range(7)
QP_STATE_NAMES = ["RESET", "INIT", "RTR", "RTS", "SQD", "SQE", "ERR"]


def ctrl_seg(opcode: int, wqe_idx: int, qpn: int, ds: int, fm_ce_se: int = 0, imm: int = 0, opmod: int = 0) -> bytes:
    w0 = ((opmod & 0xFF) << 24) | ((wqe_idx & 0xFFFF) << 8) | (opcode & 0xFF)
    w1 = ((qpn & 0xFFFFFF) << 8) | (ds & 0x3F)
    return struct.pack(">IIBBBBI", w0, w1, 0, 0, 0, fm_ce_se & 0xFF, imm & 0xFFFFFFFF)


def raddr_seg(raddr: int, rkey: int) -> bytes:
    return struct.pack(">QII", raddr, rkey, 0)


def data_seg(addr: int, lkey: int, nbytes: int) -> bytes:
    return struct.pack(">IIQ", nbytes & 0x7FFFFFFF, lkey, addr)


def rdma_wqe(opcode, wqe_idx, qpn, laddr, lkey, raddr, rkey, nbytes, fm_ce_se=CTRL_CQ_UPDATE, imm=0) -> bytes:
    """One WQEBB: ctrl + raddr + data, zero padded to 64 bytes (ds = 3)."""
    b = ctrl_seg(opcode, wqe_idx, qpn, 3, fm_ce_se, imm) + raddr_seg(raddr, rkey) + data_seg(laddr, lkey, nbytes)
    return b + bytes(WQEBB - len(b))


def send_wqe(opcode, wqe_idx, qpn, laddr, lkey, nbytes, fm_ce_se=CTRL_CQ_UPDATE, imm=0) -> bytes:
    b = ctrl_seg(opcode, wqe_idx, qpn, 2, fm_ce_se, imm) + data_seg(laddr, lkey, nbytes)
    return b + bytes(WQEBB - len(b))


@dataclass
class WqeView:
    opcode: int
    opmod: int
    wqe_idx: int
    qpn: int
    ds: int
    fm_ce_se: int
    imm: int
    raddr: int = 0
    rkey: int = 0
    laddr: int = 0
    lkey: int = 0
    nbytes: int = 0


def decode_wqe(b: bytes) -> WqeView:
    assert len(b) >= WQEBB
    w0, w1, _sig, _r0, _r1, fm, imm = struct.unpack_from(">IIBBBBI", b, 0)
    v = WqeView(opcode=w0 & 0xFF, opmod=w0 >> 24, wqe_idx=(w0 >> 8) & 0xFFFF, qpn=w1 >> 8, ds=w1 & 0x3F,
                fm_ce_se=fm, imm=imm)
    if v.opcode in (OP_RDMA_WRITE, OP_RDMA_WRITE_IMM, OP_RDMA_READ):
        v.raddr, v.rkey, _ = struct.unpack_from(">QII", b, 16)
        n, v.lkey, v.laddr = struct.unpack_from(">IIQ", b, 32)
        v.nbytes = n & 0x7FFFFFFF
    elif v.opcode in (OP_SEND, OP_SEND_IMM):
        n, v.lkey, v.laddr = struct.unpack_from(">IIQ", b, 16)
        v.nbytes = n & 0x7FFFFFFF
    return v


@dataclass
class CqeView:
    opcode: int
    owner: int
    wqe_counter: int
    qpn: int
    wqe_opcode: int
    byte_cnt: int
    imm: int
    syndrome: int
    vendor_synd: int

    @property
    def is_error(self) -> bool:
        return self.opcode in (CQE_REQ_ERR, CQE_RESP_ERR)


def cqe(opcode, owner, wqe_counter, qpn, wqe_opcode=0, byte_cnt=0, imm=0, syndrome=0, timestamp=0) -> bytes:
    b = bytearray(CQE_SIZE)
    err = opcode in (CQE_REQ_ERR, CQE_RESP_ERR)
    if not err:
        struct.pack_into(">I", b, 36, imm)
        struct.pack_into(">I", b, 44, byte_cnt)
        struct.pack_into(">II", b, 48, (timestamp >> 32) & 0xFFFFFFFF, timestamp & 0xFFFFFFFF)
    else:
        b[54] = 0
        b[55] = syndrome
    struct.pack_into(">I", b, 56, ((wqe_opcode & 0xFF) << 24) | (qpn & 0xFFFFFF))
    struct.pack_into(">H", b, 60, wqe_counter & 0xFFFF)
    b[63] = ((opcode & 0xF) << 4) | (owner & 1)
    return bytes(b)


def decode_cqe(b: bytes) -> CqeView:
    assert len(b) >= CQE_SIZE
    op_own = b[63]
    opcode = op_own >> 4
    sq, = struct.unpack_from(">I", b, 56)
    wc, = struct.unpack_from(">H", b, 60)
    err = opcode in (CQE_REQ_ERR, CQE_RESP_ERR)
    return CqeView(opcode=opcode, owner=op_own & 1, wqe_counter=wc, qpn=sq & 0xFFFFFF, wqe_opcode=sq >> 24,
                   byte_cnt=0 if err else struct.unpack_from(">I", b, 44)[0],
                   imm=0 if err else struct.unpack_from(">I", b, 36)[0],
                   syndrome=b[55] if err else 0, vendor_synd=b[54] if err else 0)


def cqe_valid(op_own: int, ci: int, log_n: int) -> bool:
    return (op_own >> 4) != CQE_INVALID and (op_own & 1) == ((ci >> log_n) & 1)


def doorbell_value(wqe_idx: int, qpn: int) -> int:
    """The 8 bytes stored to the doorbell register, as the little-endian u64 the GPU writes."""
    lo = struct.unpack("<I", struct.pack(">I", ((wqe_idx & 0xFFFF) << 8) | OP_NOP))[0]
    hi = struct.unpack("<I", struct.pack(">I", (qpn & 0xFFFFFF) << 8))[0]
    return lo | (hi << 32)


def expand16(counter16: int, near: int) -> int:
    """Expand a 16-bit wire counter to the 64-bit index closest above ``near``."""
    return near + ((counter16 - near) & 0xFFFF)
