// mlx5 work-queue wire format: WQE segments, CQE64, doorbell record.
//
// Everything a ConnectX HCA reads or writes in a send queue / completion queue,
// byte-exact and big-endian, usable from host C++, CUDA device code and (through
// the C API) the Python unit tests.  The same encoders feed a real mlx5 QP (verbs
// backend) and the software HCA (softhca backend), so the device-side posting code
// is identical on both wires.
//
// Role in the reference: none of this exists there -- amdp2p only makes
// ibv_reg_mr() succeed (amdp2p.c:363-371) and leaves posting to "IB verbs"
// (README.md:67).  SURVEY.md N3/K1 asks for the GPU to build these itself.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RN_HD __host__ __device__ __forceinline__
#else
#define RN_HD inline
#endif

namespace rn {

// ---------------------------------------------------------------- byte order
RN_HD uint32_t bswap32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(v, 0, 0x0123);
#else
  return __builtin_bswap32(v);
#endif
}
RN_HD uint16_t bswap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }
RN_HD uint64_t bswap64(uint64_t v) {
  return ((uint64_t)bswap32((uint32_t)v) << 32) | bswap32((uint32_t)(v >> 32));
}
// The build only targets little-endian hosts (x86-64) and GPUs.
RN_HD uint32_t be32(uint32_t v) { return bswap32(v); }
RN_HD uint16_t be16(uint16_t v) { return bswap16(v); }
RN_HD uint64_t be64(uint64_t v) { return bswap64(v); }

// ---------------------------------------------------------------- constants
enum : uint32_t {
  WQEBB = 64,        // basic block of the send queue
  WQE_SEG = 16,      // every segment is 16 bytes; ctrl.ds counts these
  CQE_SIZE = 64,
};

enum Opcode : uint8_t {  // ctrl.opmod_idx_opcode[7:0]
  OP_NOP = 0x00,
  OP_SEND_INVAL = 0x01,
  OP_RDMA_WRITE = 0x08,
  OP_RDMA_WRITE_IMM = 0x09,
  OP_SEND = 0x0a,
  OP_SEND_IMM = 0x0b,
  OP_RDMA_READ = 0x10,
  OP_ATOMIC_CS = 0x11,
  OP_ATOMIC_FA = 0x12,
};

enum CtrlFlags : uint8_t {  // ctrl.fm_ce_se
  CTRL_SOLICITED = 1 << 1,
  CTRL_CQ_UPDATE = 2 << 2,     // 0x08: generate a CQE for this WQE
  CTRL_FENCE = 4 << 5,         // 0x80: wait for prior reads before starting
  CTRL_INITIATOR_SMALL_FENCE = 1 << 5,
};

enum CqeOpcode : uint8_t {  // cqe.op_own[7:4]
  CQE_REQ = 0x0,
  CQE_RESP_WR_IMM = 0x1,
  CQE_RESP_SEND = 0x2,
  CQE_RESP_SEND_IMM = 0x3,
  CQE_RESP_SEND_INV = 0x4,
  CQE_RESIZE_CQ = 0x5,
  CQE_REQ_ERR = 0xd,
  CQE_RESP_ERR = 0xe,
  CQE_INVALID = 0xf,
};

enum CqeSyndrome : uint8_t {  // err_cqe.syndrome
  SYN_OK = 0x00,
  SYN_LOCAL_LENGTH_ERR = 0x01,
  SYN_LOCAL_QP_OP_ERR = 0x02,
  SYN_LOCAL_PROT_ERR = 0x04,
  SYN_WR_FLUSH_ERR = 0x05,
  SYN_MW_BIND_ERR = 0x06,
  SYN_BAD_RESP_ERR = 0x10,
  SYN_LOCAL_ACCESS_ERR = 0x11,
  SYN_REMOTE_INVAL_REQ_ERR = 0x12,
  SYN_REMOTE_ACCESS_ERR = 0x13,
  SYN_REMOTE_OP_ERR = 0x14,
  SYN_TRANSPORT_RETRY_EXC_ERR = 0x15,
  SYN_RNR_RETRY_EXC_ERR = 0x16,
  SYN_REMOTE_ABORTED_ERR = 0x22,
};

enum DbrIndex : uint32_t { DBR_RCV = 0, DBR_SND = 1 };

// ---------------------------------------------------------------- segments
struct alignas(16) CtrlSeg {     // mlx5_wqe_ctrl_seg
  uint32_t opmod_idx_opcode;     // be32: opmod[31:24] | wqe_index[23:8] | opcode[7:0]
  uint32_t qpn_ds;               // be32: qpn[31:8] | ds[5:0]  (ds = #16-byte segments)
  uint8_t signature;
  uint8_t rsvd[2];
  uint8_t fm_ce_se;
  uint32_t imm;                  // be32 immediate / invalidation key
};
struct alignas(16) RaddrSeg {    // mlx5_wqe_raddr_seg
  uint64_t raddr;                // be64
  uint32_t rkey;                 // be32
  uint32_t reserved;
};
struct alignas(16) DataSeg {     // mlx5_wqe_data_seg
  uint32_t byte_count;           // be32 (bit 31 = inline)
  uint32_t lkey;                 // be32
  uint64_t addr;                 // be64
};
struct alignas(16) AtomicSeg {   // mlx5_wqe_atomic_seg
  uint64_t swap_add;             // be64
  uint64_t compare;              // be64
};
// One WQEBB holding ctrl + raddr + one data segment (RDMA write / read), or
// ctrl + data (send: raddr slot left zero and ds=2 with data in slot 1).
struct alignas(64) Wqe64 {
  CtrlSeg ctrl;
  union {
    struct { RaddrSeg raddr; DataSeg data; uint8_t pad[16]; } rdma;
    struct { DataSeg data; uint8_t pad[32]; } send;
    struct { RaddrSeg raddr; AtomicSeg atomic; DataSeg data; } atom;
  };
};
static_assert(sizeof(CtrlSeg) == 16, "ctrl seg");
static_assert(sizeof(RaddrSeg) == 16, "raddr seg");
static_assert(sizeof(DataSeg) == 16, "data seg");
static_assert(sizeof(Wqe64) == 64, "wqebb");

// Receive WQE: a bare scatter list; we use one data segment per 16-byte stride.
struct alignas(16) RecvWqe { DataSeg data; };
static_assert(sizeof(RecvWqe) == 16, "recv wqe");

struct alignas(64) Cqe64 {       // mlx5_cqe64
  uint8_t outer_l3_tunneled;
  uint8_t rsvd0;
  uint16_t wqe_id;               // be16
  uint8_t lro_tcppsh_abort_dupack;
  uint8_t lro_min_ttl;
  uint16_t lro_tcp_win;
  uint32_t lro_ack_seq_num;
  uint32_t rss_hash_result;
  uint8_t rss_hash_type;
  uint8_t ml_path;
  uint8_t rsvd20[2];
  uint16_t check_sum;
  uint16_t slid;
  uint32_t flags_rqpn;           // be32
  uint8_t hds_ip_ext;
  uint8_t l4_l3_hdr_type;
  uint16_t vlan_info;
  uint32_t srqn;                 // be32
  uint32_t imm_inval_pkey;       // be32
  uint8_t rsvd40[4];
  uint32_t byte_cnt;             // be32
  uint32_t timestamp_h;          // be32
  uint32_t timestamp_l;          // be32
  uint32_t sop_drop_qpn;         // be32: send opcode[31:24] | qpn[23:0]
  uint16_t wqe_counter;          // be16
  uint8_t signature;
  uint8_t op_own;                // opcode[7:4] | se[1] | owner[0]
};
struct alignas(64) ErrCqe {      // mlx5_err_cqe (same 64 bytes, error view)
  uint8_t rsvd0[32];
  uint32_t srqn;
  uint8_t rsvd1[18];
  uint8_t vendor_err_synd;
  uint8_t syndrome;
  uint32_t s_wqe_opcode_qpn;     // be32: wqe opcode[31:24] | qpn[23:0]
  uint16_t wqe_counter;          // be16
  uint8_t signature;
  uint8_t op_own;
};
static_assert(sizeof(Cqe64) == 64, "cqe64");
static_assert(sizeof(ErrCqe) == 64, "err cqe");

// ---------------------------------------------------------------- encoders
RN_HD uint32_t ctrl_word0(uint8_t opcode, uint16_t wqe_idx, uint8_t opmod = 0) {
  return be32(((uint32_t)opmod << 24) | ((uint32_t)wqe_idx << 8) | opcode);
}
RN_HD uint32_t ctrl_word1(uint32_t qpn, uint8_t ds) { return be32((qpn << 8) | (ds & 0x3f)); }

RN_HD void encode_ctrl(CtrlSeg* c, uint8_t opcode, uint16_t wqe_idx, uint32_t qpn, uint8_t ds,
                       uint8_t fm_ce_se, uint32_t imm) {
  c->opmod_idx_opcode = ctrl_word0(opcode, wqe_idx);
  c->qpn_ds = ctrl_word1(qpn, ds);
  c->signature = 0;
  c->rsvd[0] = c->rsvd[1] = 0;
  c->fm_ce_se = fm_ce_se;
  c->imm = be32(imm);
}
RN_HD void encode_raddr(RaddrSeg* r, uint64_t raddr, uint32_t rkey) {
  r->raddr = be64(raddr);
  r->rkey = be32(rkey);
  r->reserved = 0;
}
RN_HD void encode_data(DataSeg* d, uint64_t addr, uint32_t lkey, uint32_t bytes) {
  d->byte_count = be32(bytes & 0x7fffffffu);
  d->lkey = be32(lkey);
  d->addr = be64(addr);
}

// Full WQE builders (host and slow-path device use; the fast device path in
// hca/post.cuh emits the same bytes with four 16-byte vector stores).
RN_HD void build_rdma_wqe(Wqe64* w, uint8_t opcode, uint16_t wqe_idx, uint32_t qpn, uint64_t laddr,
                          uint32_t lkey, uint64_t raddr, uint32_t rkey, uint32_t bytes,
                          uint8_t fm_ce_se, uint32_t imm) {
  encode_ctrl(&w->ctrl, opcode, wqe_idx, qpn, 3, fm_ce_se, imm);
  encode_raddr(&w->rdma.raddr, raddr, rkey);
  encode_data(&w->rdma.data, laddr, lkey, bytes);
  for (int i = 0; i < 16; ++i) w->rdma.pad[i] = 0;
}
RN_HD void build_send_wqe(Wqe64* w, uint8_t opcode, uint16_t wqe_idx, uint32_t qpn, uint64_t laddr,
                          uint32_t lkey, uint32_t bytes, uint8_t fm_ce_se, uint32_t imm) {
  encode_ctrl(&w->ctrl, opcode, wqe_idx, qpn, 2, fm_ce_se, imm);
  encode_data(&w->send.data, laddr, lkey, bytes);
  for (int i = 0; i < 32; ++i) w->send.pad[i] = 0;
}

// ---------------------------------------------------------------- decoders
struct WqeView {
  uint8_t opcode, opmod, ds, fm_ce_se;
  uint16_t wqe_idx;
  uint32_t qpn, imm;
  uint64_t raddr, laddr;
  uint32_t rkey, lkey, bytes;
};
RN_HD bool decode_wqe(const Wqe64* w, WqeView* v) {
  uint32_t w0 = be32(w->ctrl.opmod_idx_opcode), w1 = be32(w->ctrl.qpn_ds);
  v->opcode = (uint8_t)(w0 & 0xff);
  v->wqe_idx = (uint16_t)((w0 >> 8) & 0xffff);
  v->opmod = (uint8_t)(w0 >> 24);
  v->qpn = w1 >> 8;
  v->ds = (uint8_t)(w1 & 0x3f);
  v->fm_ce_se = w->ctrl.fm_ce_se;
  v->imm = be32(w->ctrl.imm);
  v->raddr = 0; v->rkey = 0; v->laddr = 0; v->lkey = 0; v->bytes = 0;
  switch (v->opcode) {
    case OP_RDMA_WRITE: case OP_RDMA_WRITE_IMM: case OP_RDMA_READ:
      if (v->ds != 3) return false;
      v->raddr = be64(w->rdma.raddr.raddr);
      v->rkey = be32(w->rdma.raddr.rkey);
      v->laddr = be64(w->rdma.data.addr);
      v->lkey = be32(w->rdma.data.lkey);
      v->bytes = be32(w->rdma.data.byte_count) & 0x7fffffffu;
      return true;
    case OP_SEND: case OP_SEND_IMM:
      if (v->ds != 2) return false;
      v->laddr = be64(w->send.data.addr);
      v->lkey = be32(w->send.data.lkey);
      v->bytes = be32(w->send.data.byte_count) & 0x7fffffffu;
      return true;
    case OP_NOP:
      return v->ds == 1;
    default:
      return false;
  }
}

RN_HD uint8_t cqe_op_own(uint8_t opcode, uint8_t owner) { return (uint8_t)((opcode << 4) | (owner & 1)); }
RN_HD uint8_t cqe_opcode(uint8_t op_own) { return op_own >> 4; }
RN_HD uint8_t cqe_owner(uint8_t op_own) { return op_own & 1; }

struct CqeView {
  uint8_t opcode, owner, syndrome, vendor_synd, wqe_opcode;
  uint16_t wqe_counter;
  uint32_t qpn, byte_cnt, imm;
  bool is_error;
};
RN_HD void decode_cqe(const Cqe64* c, CqeView* v) {
  v->opcode = cqe_opcode(c->op_own);
  v->owner = cqe_owner(c->op_own);
  v->wqe_counter = be16(c->wqe_counter);
  uint32_t sq = be32(c->sop_drop_qpn);
  v->qpn = sq & 0xffffff;
  v->wqe_opcode = (uint8_t)(sq >> 24);
  v->byte_cnt = be32(c->byte_cnt);
  v->imm = be32(c->imm_inval_pkey);
  v->is_error = (v->opcode == CQE_REQ_ERR || v->opcode == CQE_RESP_ERR);
  const ErrCqe* e = reinterpret_cast<const ErrCqe*>(c);
  v->syndrome = v->is_error ? e->syndrome : 0;
  v->vendor_synd = v->is_error ? e->vendor_err_synd : 0;
}

// A CQE at consumer index `ci` of a 2^log_n ring is valid when its opcode is not
// INVALID and its owner bit equals the pass parity of ci.
RN_HD bool cqe_valid(uint8_t op_own, uint32_t ci, uint32_t log_n) {
  return cqe_opcode(op_own) != CQE_INVALID && cqe_owner(op_own) == ((ci >> log_n) & 1u);
}

}  // namespace rn
