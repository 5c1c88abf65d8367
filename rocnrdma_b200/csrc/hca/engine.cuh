// The softhca DMA engine: a persistent sm_100a kernel that plays the HCA.
//
// Each engine CTA scans the doorbell registers of the QP table, claims either the
// prologue of the next un-parsed WQE (decode, MKey checks, address translation,
// receive-WQE matching) or one chunk of an already-parsed WQE, moves the chunk
// with a TMA bulk-copy pipeline (cp.async.bulk global->shared->global, one
// issuing thread, mbarrier-tracked, no LSU traffic), and retires WQEs strictly in
// order by writing mlx5 CQEs.  Messages therefore fan out over every engine CTA
// (large transfers run at HBM / NVLink speed) while small messages cost one
// doorbell poll + one prologue.
//
// Failure handling mirrors an RC QP: a bad key / bounds / opcode produces an error
// CQE with an IB syndrome, moves the QP to ERR and flushes later WQEs; a missing
// receive WQE is retried until rnr_timeout_ns, then fails with RNR_RETRY_EXC.
// Nothing spins unbounded: the engine leaves when *stop is set or when no doorbell
// has moved for idle_timeout_ns (watchdog), so a forgotten engine cannot wedge a
// GPU ("bounded spin + status word").
//
// Reference parity: this is the role the Mellanox HCA itself plays below
// amdp2p's dma_map (amdp2p.c:219-264): consuming bus addresses and moving bytes.
#pragma once
#include <cuda_runtime.h>
#include "hca_types.h"
#include "post.cuh"

namespace rn {
namespace eng {

using namespace rn::dev;

constexpr int kThreads = 128;
constexpr uint32_t kSub = 16384;        // bytes per TMA bulk transaction
constexpr int kStages = 12;             // smem ring depth (12 x 16 KiB = 192 KiB in flight per SM)
constexpr uint32_t kBulkMin = 4096;     // below this the generic path is as fast

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ MKey lookup
// Returns the engine-reachable pointer for [addr, addr+len) under `key`, or 0 with
// *syn set.  need = access bits that must all be present.
__device__ __forceinline__ uint64_t translate(const MKeyEntry* tab, uint32_t n, uint32_t key, uint64_t addr,
                                              uint32_t len, uint32_t need, bool remote, uint8_t* syn) {
  uint32_t idx = key >> 8;
  if (idx >= n) { *syn = remote ? SYN_REMOTE_ACCESS_ERR : SYN_LOCAL_PROT_ERR; return 0; }
  const MKeyEntry* e = tab + idx;
  uint4 a = ld_v4_volatile(e);                                   // base, len
  uint4 b = ld_v4_volatile(reinterpret_cast<const uint8_t*>(e) + 16);  // map_base, key, access
  uint4 c = ld_v4_volatile(reinterpret_cast<const uint8_t*>(e) + 32);             // valid, kind
  uint64_t base = ((uint64_t)a.y << 32) | a.x, mlen = ((uint64_t)a.w << 32) | a.z;
  uint64_t map_base = ((uint64_t)b.y << 32) | b.x;
  uint32_t ekey = b.z, acc = b.w, valid = c.x;
  if (!valid || ekey != key) { *syn = remote ? SYN_REMOTE_ACCESS_ERR : SYN_LOCAL_PROT_ERR; return 0; }
  if ((acc & need) != need) { *syn = remote ? SYN_REMOTE_ACCESS_ERR : SYN_LOCAL_ACCESS_ERR; return 0; }
  if (addr < base || addr + len > base + mlen || addr + len < addr) {
    *syn = remote ? SYN_REMOTE_ACCESS_ERR : SYN_LOCAL_PROT_ERR;
    return 0;
  }
  if (len == 0) return map_base ? map_base : 1;  // zero-length: any non-null token
  return map_base + (addr - base);
}

// ------------------------------------------------------------------ CQE writer
__device__ __forceinline__ void write_cqe(CqDev* cq, uint8_t* ring, uint8_t opcode, uint8_t wqe_opcode,
                                          uint32_t qpn, uint16_t wqe_counter, uint32_t byte_cnt, uint32_t imm,
                                          uint8_t syndrome, bool sys) {
  unsigned int slot = sys ? atomicAdd_system(&cq->pi, 1u) : atomicAdd(&cq->pi, 1u);
  uint32_t log_n = cq->log_n;
  uint8_t* cqe = ring + ((size_t)(slot & ((1u << log_n) - 1)) << 6);
  uint8_t owner = (uint8_t)((slot >> log_n) & 1u);
  unsigned long long now = globaltimer_ns();
  bool err = (opcode == CQE_REQ_ERR || opcode == CQE_RESP_ERR);
  st_v4(cqe + 0, 0u, 0u, 0u, 0u);
  st_v4(cqe + 16, 0u, 0u, 0u, 0u);
  if (!err) {
    st_v4(cqe + 32, 0u, be32(imm), 0u, be32(byte_cnt));
  } else {
    // err view: bytes 54 = vendor_err_synd, 55 = syndrome  (word at 52..55)
    st_v4(cqe + 32, 0u, 0u, 0u, 0u);
  }
  uint32_t w48 = err ? 0u : be32((uint32_t)(now >> 32));
  uint32_t w52 = err ? ((uint32_t)syndrome << 24) : be32((uint32_t)now);
  uint32_t w56 = be32(((uint32_t)wqe_opcode << 24) | (qpn & 0xffffff));
  uint32_t w60 = (uint32_t)be16(wqe_counter) | ((uint32_t)cqe_op_own(opcode, owner) << 24);
  fence_scope(sys);  // payload + first 48 bytes before the word that flips ownership
  st_v4(cqe + 48, w48, w52, w56, w60);
}

// ------------------------------------------------------------------ chunk copy
struct Smem {
  alignas(128) uint8_t ring[kStages][kSub];
  alignas(8) uint64_t full[kStages];
  // work descriptor broadcast from thread 0
  uint64_t src, dst;
  uint32_t len;
  int have_work;
  uint32_t phase_bits;   // per-stage parity of the next wait
};

// Bulk path: thread 0 runs loads kStages-1 ahead of stores.  Requires 16-byte
// aligned src, dst and len.
__device__ __forceinline__ bool copy_bulk(Smem& s, uint64_t src, uint64_t dst, uint32_t len) {
  const uint32_t nsub = (len + kSub - 1) / kSub;
  uint32_t phase_bits = s.phase_bits;
  constexpr int P = kStages - 1;
  auto sub_len = [&](uint32_t i) { return (i + 1 == nsub) ? (len - i * kSub) : kSub; };
  uint32_t issued = 0;
  for (; issued < nsub && issued < (uint32_t)P; ++issued) {
    int st = issued % kStages;
    mbar_expect_tx(&s.full[st], sub_len(issued));
    bulk_g2s(s.ring[st], (const void*)(src + (uint64_t)issued * kSub), sub_len(issued), &s.full[st]);
  }
  for (uint32_t i = 0; i < nsub; ++i) {
    int st = i % kStages;
    if (!mbar_try_wait(&s.full[st], (phase_bits >> st) & 1u)) {
      // bounded: a bulk load that never lands (bad mapping) must not wedge the SM
      unsigned long long t0 = globaltimer_ns();
      while (!mbar_try_wait(&s.full[st], (phase_bits >> st) & 1u)) {
        if (globaltimer_ns() - t0 > 1000000000ull) { s.phase_bits = phase_bits; return false; }
      }
    }
    phase_bits ^= 1u << st;
    bulk_s2g((void*)(dst + (uint64_t)i * kSub), s.ring[st], sub_len(i));
    bulk_commit();
    if (issued < nsub) {
      // stage of load `issued` was last read by store i-1: allow only store i to be pending
      bulk_wait_read<1>();
      int ls = issued % kStages;
      mbar_expect_tx(&s.full[ls], sub_len(issued));
      bulk_g2s(s.ring[ls], (const void*)(src + (uint64_t)issued * kSub), sub_len(issued), &s.full[ls]);
      ++issued;
    }
  }
  bulk_wait_all();  // writes complete and visible to this thread
  s.phase_bits = phase_bits;
  return true;
}

// Generic path: any alignment, whole CTA.
__device__ __forceinline__ void copy_generic(uint64_t src, uint64_t dst, uint32_t len) {
  const uint8_t* s = (const uint8_t*)src;
  uint8_t* d = (uint8_t*)dst;
  if ((((src ^ dst) & 15) == 0) && len >= 64) {
    uint32_t head = (uint32_t)((16 - (src & 15)) & 15);
    if (head > len) head = len;
    for (uint32_t i = threadIdx.x; i < head; i += kThreads) d[i] = s[i];
    uint32_t body = (len - head) & ~15u;
    const uint4* s4 = (const uint4*)(s + head);
    uint4* d4 = (uint4*)(d + head);
    for (uint32_t i = threadIdx.x; i < body / 16; i += kThreads) d4[i] = s4[i];
    for (uint32_t i = head + body + threadIdx.x; i < len; i += kThreads) d[i] = s[i];
  } else {
    for (uint32_t i = threadIdx.x; i < len; i += kThreads) d[i] = s[i];
  }
}

// ------------------------------------------------------------------ prologue
// Parse WQE `w` of `qp`, fill resolved[slot].  Returns false when the WQE must be
// retried later (receiver not ready).
__device__ __forceinline__ bool prologue(EngineCtl* ctl, QpDev* qp, unsigned long long w) {
  const uint32_t mask = (1u << qp->sq_log) - 1;
  Resolved* r = qp->resolved + (w & mask);
  const uint8_t* slot = qp->sq + ((w & mask) << 6);
  Wqe64 wqe;
  uint4* wv = reinterpret_cast<uint4*>(&wqe);
  wv[0] = ld_v4_volatile(slot);
  wv[1] = ld_v4_volatile(slot + 16);
  wv[2] = ld_v4_volatile(slot + 32);
  wv[3] = ld_v4_volatile(slot + 48);
  WqeView v;
  uint8_t syn = SYN_OK;
  uint64_t src = 0, dst = 0;
  uint64_t rq_idx = 0;
  uint8_t rq_taken = 0;
  bool ok = decode_wqe(&wqe, &v);
  if (qp->state == QPS_ERR) {
    syn = SYN_WR_FLUSH_ERR;
  } else if (!ok || v.qpn != qp->qpn || v.wqe_idx != (uint16_t)w) {
    syn = SYN_LOCAL_QP_OP_ERR;
  } else if (!qp->r.connected && v.opcode != OP_NOP) {
    syn = SYN_LOCAL_QP_OP_ERR;
  } else {
    switch (v.opcode) {
      case OP_NOP: v.bytes = 0; break;
      case OP_RDMA_WRITE:
      case OP_RDMA_WRITE_IMM:
        src = translate(qp->lkeys, qp->n_lkeys, v.lkey, v.laddr, v.bytes, 0, false, &syn);
        if (src) dst = translate(qp->r.rkeys, qp->r.n_rkeys, v.rkey, v.raddr, v.bytes, ACC_REMOTE_WRITE, true, &syn);
        break;
      case OP_RDMA_READ:
        dst = translate(qp->lkeys, qp->n_lkeys, v.lkey, v.laddr, v.bytes, ACC_LOCAL_WRITE, false, &syn);
        if (dst) src = translate(qp->r.rkeys, qp->r.n_rkeys, v.rkey, v.raddr, v.bytes, ACC_REMOTE_READ, true, &syn);
        break;
      case OP_SEND:
      case OP_SEND_IMM:
        src = translate(qp->lkeys, qp->n_lkeys, v.lkey, v.laddr, v.bytes, 0, false, &syn);
        break;
      default: syn = SYN_LOCAL_QP_OP_ERR;
    }
    bool needs_recv = (syn == SYN_OK) &&
                      (v.opcode == OP_SEND || v.opcode == OP_SEND_IMM || v.opcode == OP_RDMA_WRITE_IMM);
    if (needs_recv) {
      // match against the responder's receive queue
      uint32_t rpi16 = be32(ld_u32_volatile(&qp->r.rq_dbr[DBR_RCV])) & 0xffff;
      unsigned long long head = qp->rq_head;
      if (((rpi16 - (uint32_t)head) & 0xffff) == 0) {
        unsigned long long now = globaltimer_ns();
        if (qp->rnr_since == 0) { qp->rnr_since = now; atomicAdd(&qp->n_rnr, 1ull); }
        if (now - qp->rnr_since < ctl->rnr_timeout_ns) return false;  // retry later
        syn = SYN_RNR_RETRY_EXC_ERR;
      } else {
        fence_scope(qp->sys_scope != 0);
        const uint8_t* rs = qp->r.rq + ((head & ((1ull << qp->r.rq_log) - 1)) << 4);
        uint4 d = ld_v4_volatile(rs);
        uint32_t rbytes = be32(d.x) & 0x7fffffffu, rlkey = be32(d.y);
        uint64_t raddr = ((uint64_t)be32(d.z) << 32) | be32(d.w);
        rq_idx = head;
        rq_taken = 1;
        qp->rq_head = head + 1;
        if (v.opcode != OP_RDMA_WRITE_IMM) {
          if (rbytes < v.bytes) syn = SYN_REMOTE_INVAL_REQ_ERR;
          else dst = translate(qp->r.rkeys, qp->r.n_rkeys, rlkey, raddr, v.bytes, ACC_LOCAL_WRITE, true, &syn);
        }
      }
      qp->rnr_since = 0;
    }
  }
  // Work granule: at least the QP's chunk_bytes, grown so that a large message is cut
  // into ~4 claims per engine CTA.  Claims serialise on one atomic (~1 us each), so a
  // fixed 128 KiB granule caps a 1 GiB write near 130 GB/s (measured); this keeps the
  // claim rate negligible while still load-balancing the tail.
  uint32_t chunk = qp->chunk_bytes;
  {
    uint32_t target = 8u * gridDim.x;
    uint32_t want = (uint32_t)(((uint64_t)v.bytes + target - 1) / target);
    want = (want + kSub - 1) / kSub * kSub;
    if (want > chunk) chunk = want;
  }
  uint32_t nchunks = (syn == SYN_OK && v.bytes > 0) ? (v.bytes + chunk - 1) / chunk : 1;
  r->chunk = chunk;
  r->src = src; r->dst = dst;
  r->bytes = (syn == SYN_OK) ? v.bytes : 0;
  r->nchunks = nchunks;
  r->imm = v.imm;
  r->opcode = v.opcode; r->fm_ce_se = v.fm_ce_se; r->syndrome = syn; r->rq_consumed = rq_taken;
  r->rq_idx = rq_idx;
  r->done = 0;
  *(volatile unsigned long long*)&r->state = (w << 2) | 1ull;
  if (syn != SYN_OK && syn != SYN_WR_FLUSH_ERR) qp->state = QPS_ERR;
  atomicAdd(&qp->n_wqe, 1ull);
  trace_stamp(qp, w, TR_PARSED);
  return true;
}

// ------------------------------------------------------------------ retire
__device__ __forceinline__ void retire(QpDev* qp) {
  const uint32_t mask = (1u << qp->sq_log) - 1;
  const bool sys = qp->sys_scope != 0;
  for (;;) {
    if (atomicCAS(&qp->retire_lock, 0u, 1u) != 0u) return;
    fence_gpu();  // acquire: retire_head and slot states written by the previous holder / finishers
    unsigned long long h = ld_u64_volatile(&qp->retire_head);
    for (;;) {
      Resolved* r = qp->resolved + (h & mask);
      if (ld_u64_volatile(&r->state) != ((h << 2) | 2ull)) break;
      uint8_t opc = r->opcode, syn = r->syndrome;
      bool err = syn != SYN_OK;
      if (r->rq_consumed && qp->r.rcq) {
        uint8_t ropc = err ? CQE_RESP_ERR
                           : (opc == OP_SEND ? CQE_RESP_SEND
                                             : (opc == OP_SEND_IMM ? CQE_RESP_SEND_IMM : CQE_RESP_WR_IMM));
        write_cqe(qp->r.rcq, qp->r.rcq_buf, ropc, 0, qp->r.qpn, (uint16_t)r->rq_idx, r->bytes, r->imm, syn, sys);
      }
      if (err || (r->fm_ce_se & CTRL_CQ_UPDATE)) {
        write_cqe(qp->scq, qp->scq->buf, err ? CQE_REQ_ERR : CQE_REQ, opc, qp->qpn, (uint16_t)h, r->bytes, 0, syn, sys);
        trace_stamp(qp, h, TR_CQE);
        qp->n_cqe = qp->n_cqe + 1;          // counters are only written under the retire lock
      }
      if (err) qp->n_err = qp->n_err + 1;
      qp->n_bytes = qp->n_bytes + r->bytes;
      ++h;
    }
    *(volatile unsigned long long*)&qp->retire_head = h;
    fence_gpu();  // release
    atomicExch(&qp->retire_lock, 0u);
    Resolved* r = qp->resolved + (h & mask);
    if (ld_u64_volatile(&r->state) != ((h << 2) | 2ull)) return;
  }
}

// ------------------------------------------------------------------ claim
struct Work {
  QpDev* qp;
  unsigned long long w;
  uint32_t chunk;
};

__device__ __forceinline__ bool try_claim(EngineCtl* ctl, QpDev* qp, Work* out, bool* saw_pending) {
  if (qp->state != QPS_RTS && qp->state != QPS_ERR) return false;
  const bool sys = qp->sys_scope != 0;
  unsigned long long cur = ld_u64_volatile(&qp->cursor);
  unsigned long long w = cur >> CURSOR_CHUNK_BITS;
  const uint32_t ph = (uint32_t)(cur & CURSOR_PHASE_MASK);
  if (ph == PH_LOCKED) { *saw_pending = true; return false; }
  const uint32_t mask = (1u << qp->sq_log) - 1;
  if (ph == PH_UNPARSED) {
    unsigned long long db = ld_u64_volatile(qp->bf);
    uint32_t idx16 = (be32((uint32_t)db) >> 8) & 0xffff;
    uint32_t pending = (idx16 + 1 - (uint32_t)w) & 0xffff;
    if (pending == 0) return false;
    *saw_pending = true;
    if (atomicCAS(&qp->cursor, cur, (w << CURSOR_CHUNK_BITS) | PH_LOCKED) != cur) return false;
    fence_scope(sys);  // acquire: WQE bytes the doorbell announced
    trace_stamp(qp, w, TR_CLAIM);
    // ordering audit: the doorbell record must already cover what the register announced
    uint32_t dbr16 = be32(ld_u32_volatile(&qp->dbr[DBR_SND])) & 0xffff;
    if (((dbr16 - (uint32_t)w) & 0xffff) < pending && ((dbr16 - (uint32_t)w) & 0xffff) < 0x8000)
      atomicAdd(&qp->n_db_order_violations, 1ull);
    if (!prologue(ctl, qp, w)) {
      // receiver not ready: hand the WQE back (the one legal backwards move; nothing else
      // can touch the cursor while it is PH_LOCKED)
      atomicExch(&qp->cursor, w << CURSOR_CHUNK_BITS);
      return false;
    }
    Resolved* res = qp->resolved + (w & mask);
    uint32_t n = res->nchunks;
    if (n == 1) {
      // Single-claim WQE: nobody else reads resolved[w] before retirement, so no ticket and
      // no release fence -- just move the queue on (fire-and-forget RED, no round trip).
      atomicMax(&qp->cursor, (w + 1) << CURSOR_CHUNK_BITS);
      out->qp = qp; out->w = w; out->chunk = 0;
      return true;
    }
    fence_gpu();  // release: resolved[] fields before the ticket that lets others read them
    // Every cursor move from here on is an atomicMax, so it does not matter whether the
    // ticket or the cursor becomes visible first: a draw on a not-yet-armed ticket sees the
    // previous WQE's exhausted one, and a straggler that draws the last chunk early can only
    // push the cursor forward.
    st_u64_relaxed(&res->ticket, TICKET_ONE | ((w & TICKET_FIELD_MASK) << TICKET_GEN_SHIFT) | n);
    atomicMax(&qp->cursor, (w << CURSOR_CHUNK_BITS) | PH_OFFER);
    out->qp = qp; out->w = w; out->chunk = 0;
    return true;
  }
  // PH_OFFER: chunks of WQE w are on offer.  One fetch-add = one claim; the returned word
  // carries (chunk, generation, nchunks), so a straggler that hits a recycled slot still
  // holds a valid claim on whatever WQE the slot describes now.
  *saw_pending = true;
  Resolved* res = qp->resolved + (w & mask);
  unsigned long long t = atomicAdd(&res->ticket, (unsigned long long)TICKET_ONE);
  uint32_t c = (uint32_t)(t >> 40), gen = (uint32_t)((t >> TICKET_GEN_SHIFT) & TICKET_FIELD_MASK);
  uint32_t n = (uint32_t)(t & TICKET_FIELD_MASK);
  if (c >= n) { __nanosleep(100); return false; }
  unsigned long long wfull = w + ((gen - (uint32_t)w) & TICKET_FIELD_MASK);
  if (c + 1 == n) atomicMax(&qp->cursor, (wfull + 1) << CURSOR_CHUNK_BITS);
  fence_gpu();  // acquire: resolved[] fields of the generation this ticket names
  out->qp = qp; out->w = wfull; out->chunk = c;
  return true;
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 1) engine_kernel(EngineCtl* ctl) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  __shared__ Work work;
  __shared__ int quit;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&s.full[i], 1);
    s.phase_bits = 0;
    quit = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    atomicAdd(&ctl->running_ctas, 1u);
    if (blockIdx.x == 0) ctl->dbg_t_start = globaltimer_ns();
  }
  __syncthreads();
  unsigned long long last_activity = globaltimer_ns();
  uint32_t rr = blockIdx.x;
  unsigned spins = 0;
  for (;;) {
    if (threadIdx.x == 0) {
      s.have_work = 0;
      uint32_t n = ctl->n_qps;
      bool pending = false;
      if (blockIdx.x == 0) {
        ctl->n_polls = ctl->n_polls + 1;
        if (n) { ctl->dbg_last_db = ld_u64_volatile(ctl->qps[0]->bf); ctl->dbg_last_state = ctl->qps[0]->state; }
      }
      for (uint32_t k = 0; k < n && !s.have_work; ++k) {
        QpDev* qp = *(QpDev* volatile*)&ctl->qps[(rr + k) % n];   // table grows while we run
        if (qp && try_claim(ctl, qp, &work, &pending)) {
          s.have_work = 1;
          rr = (rr + k) % n;
        }
      }
      if (s.have_work) {
        Resolved* r = work.qp->resolved + (work.w & ((1u << work.qp->sq_log) - 1));
        uint32_t chunk = *(volatile uint32_t*)&r->chunk;
        uint64_t off = (uint64_t)work.chunk * chunk;
        uint32_t bytes = *(volatile uint32_t*)&r->bytes;
        s.len = bytes == 0 ? 0 : (uint32_t)((bytes - off < chunk) ? (bytes - off) : chunk);
        s.src = r->src + off;
        s.dst = r->dst + off;
        last_activity = globaltimer_ns();
        spins = 0;
      } else {
        ++spins;
        if (pending) last_activity = globaltimer_ns();
        if ((spins & 63) == 0) {
          fence_gpu();  // drops stale L1 lines: host-side updates (new QPs, reconnects) become visible
          if (*ctl->stop) quit = 1;
          else if (globaltimer_ns() - last_activity > ctl->idle_timeout_ns) { quit = 1; ctl->exited_idle = 1; }
        }
        if (spins > 256) __nanosleep(200);
      }
    }
    __syncthreads();
    const int q = quit, hw = s.have_work;
    const uint32_t len = s.len;
    const uint64_t csrc = s.src, cdst = s.dst;
    __syncthreads();  // every thread holds the broadcast in registers before thread 0 reuses smem
    if (q) break;
    if (!hw) continue;
    if (len > 0) {
      const bool bulk = len >= kBulkMin && (((csrc | cdst) & 15) == 0) && ((len & 15) == 0);
      if (bulk) {
        if (threadIdx.x == 0) {
          if (!copy_bulk(s, csrc, cdst, len)) { ctl->fatal = 1; quit = 1; }
          atomicAdd(&ctl->n_bulk_chunks, 1ull);
        }
      } else {
        copy_generic(csrc, cdst, len);
      }
    }
    __syncthreads();
    if (quit) break;   // fatal DMA fault: leave without completing (host sees ctl->fatal)
    if (threadIdx.x == 0) {
      QpDev* qp = work.qp;
      Resolved* r = qp->resolved + (work.w & ((1u << qp->sq_log) - 1));
      fence_scope(qp->sys_scope != 0);  // this chunk's bytes before the count that may complete the WQE
      const uint32_t nch = r->nchunks;
      if (nch == 1 || atomicAdd(&r->done, 1u) + 1 == nch) {
        trace_stamp(qp, work.w, TR_COPIED);
        fence_gpu();  // observe every other chunk's count -> their bytes precede our CQE
        *(volatile unsigned long long*)&r->state = (work.w << 2) | 2ull;
        retire(qp);
      }
    }
  }
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) ctl->dbg_t_exit = globaltimer_ns();
    atomicSub(&ctl->running_ctas, 1u);
  }
}

inline size_t engine_smem_bytes() { return sizeof(Smem) + 128; }

}  // namespace eng
}  // namespace rn
