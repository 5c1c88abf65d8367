// K4: a tcgen05 GEMM whose epilogue feeds the wire.
//
//   C[M,N] (bf16, or block-scaled fp8 panel records) = A[M,K] (bf16, K-major) * B[N,K]^T (bf16, K-major), fp32 accumulation in TMEM.
//
// Persistent, warp-specialised, one CTA per SM.  Three tile shapes share every building block in this file:
//   gemm_send_kernel   1 CTA per 128x256 tile     (gemm_tile_body<1>: cta_group::1, 4 x 48 KiB stages, TMEM double-buffered)
//   gemm_send2_kernel  CTA pair per 256x256 tile  (gemm_tile_body<2>: cta_group::2, 6 x 32 KiB stages, TMEM double-buffered)
//   gemm_send3_kernel  CTA pair per 512x256 tile  (wide: 256 rows of A per CTA, two accumulator halves, stream-K tail; own body
//                                                  because its accumulator schedule differs: see the notes above it)
// Roles:
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads (SWIZZLE_128B) into the shared-memory ring; out-of-bounds rows and
//               the K tail are zero-filled by TMA, so shapes need not be tile multiples
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::f16 (M = 128 or 256, N = 256, K = 16), four per
//               k-block per accumulator; tcgen05.commit releases smem stages and publishes finished accumulators
//   warps 2..   epilogue: tcgen05.ld 32x32b.x32 (each warp owns its 32-lane TMEM quarter), fp32 -> bf16 or -> e4m3 + UE8M0
//               block scales, staged in shared memory in the SWIZZLE_128B pattern and written with TMA tensor stores (which
//               clip at the edges) into the REGISTERED send buffer; the TMEM buffer goes back to the MMA warp as soon as its
//               last values are in registers, so tile i+1's MMAs overlap tile i's epilogue
//   send        tiles are scheduled so that 128-row panels of C complete progressively; the epilogue that finishes a panel's
//               last tile builds ONE RDMA WRITE for the panel (rows x N x 2 contiguous bytes, or one fp8 record), rings the
//               doorbell and goes back to computing -- the wire moves panel p while panels p+1.. are still being multiplied.
//               The last CTA posts a flush NOP and waits for its CQE, so the kernel's device time covers compute AND delivery.
// Every mbarrier wait is bounded (a wrong descriptor must not wedge an SM for good).
// No library GEMM anywhere on this path; the reference has no counterpart (no GPU code at all).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../hca/post.cuh"

using namespace rn;
using namespace rn::dev;

#define RN_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16, STAGES = 4;
constexpr int A_STAGE = BM * BK * 2, B_STAGE = BN * BK * 2;       // 16 KiB + 32 KiB
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr uint32_t kTmemCols = 512;                                // two 256-column accumulators
constexpr unsigned long long kWaitNs = 2000000000ull;
constexpr int kStageC = 32 * 128;                                  // one TMA-store box: 32 rows x 64 bf16, SWIZZLE_128B
constexpr uint32_t kFlagDirect = 1, kFlagPlainStores = 2, kFlagDenseProbe = 4, kFlagNoStreamK = 8;

struct alignas(1024) Smem {
  uint8_t a[STAGES][A_STAGE];
  uint8_t b[STAGES][B_STAGE];
  uint8_t stage_c[4][2][kStageC];                                  // epilogue staging: per warp, 2 x (32 rows x 128 B)
  alignas(8) uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
  volatile int abort;
};

struct GemmArgs {
  __nv_bfloat16* c;          // send buffer, row-major, ld = N (registered)
  uint32_t M, N, K;
  QpDev* qp;                 // nullptr: compute only
  uint64_t c_va;             // VA of c as registered
  uint32_t lkey, rkey;
  uint64_t remote_va;
  uint32_t signal_every;
  uint32_t with_imm;         // 1: RDMA_WRITE_IMM, immediate = panel index (wakes a consumer on the receiving GPU)
  uint32_t post_only;        // 1: post panels + flush but do not wait for the drain (profilers serialise kernels; see pack_fp8.cu)
  uint32_t out_fp8;          // 1: epilogue emits block-scaled fp8 panel records instead of bf16 rows (see below)
  uint32_t dense_probe;      // link probe ONLY (output layout is wrong on purpose): every 32x64 box lands as one contiguous 4 KiB run
  uint32_t plain_stores;     // 1: bf16 epilogue writes rows with per-thread 16-byte stores instead of staged TMA stores (A/B switch)
  uint32_t direct;           // 1: `c` IS the peer's registered buffer (NVLink-mapped): the epilogue's stores are the transfer;
                             //    each finished panel is announced by a zero-length RDMA_WRITE_IMM posted after a cumulative
                             //    system-scope fence (data plane = SM stores over NVLink, control plane = the RDMA queue pair)
  uint32_t group_m;          // tile rasterisation: this many M blocks advance together across N (L2 reuse of B)
  float* ws;                 // stream-K workspace: per cluster boundary, 2 CTAs x 2 halves x 128 x 256 fp32 partial accumulators
  unsigned int* ws_flags;    // [boundary][rank]: == epoch once that CTA's partial is complete
  uint32_t epoch;            // launch stamp (never 0): flags need no reset between launches
  unsigned int* counters;    // [0..m_blks): tiles done per panel ; [m_blks]: CTAs done
  unsigned long long* acc;   // [0] max idx+1, [1] posted, [2] ~first post time
  unsigned long long* out;   // [status, t_start, t_end, posted, t_first_post, t_compute_end, 0, 0]
  uint64_t timeout_ns;
};

// Tile order.  Plain N-fastest order re-streams the whole B matrix once per M block; at 8192^3 that is
// 4.3 GB of reads of a 134 MB operand that does not fit the 126 MB L2 -- most of the kernel's time.
// Grouped order: `group_m` M blocks advance together across N (M fastest inside the group), so the
// clusters running at the same time share B tiles in L2.  Panels of a group complete together, group
// by group, which is still progressive for the sender.
__device__ __forceinline__ void tile_coords(uint32_t t, uint32_t m_blks, uint32_t n_blks, uint32_t group_m, uint32_t* m, uint32_t* n) {
  const uint32_t per_group = group_m * n_blks;
  const uint32_t grp = t / per_group, r = t % per_group;
  const uint32_t m0 = grp * group_m;
  const uint32_t gm = min(group_m, m_blks - m0);
  *m = m0 + r % gm;
  *n = r / gm;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or ~`ns` elapse instead of
// returning at once -- a waiting role then costs no issue slots (and no power: the box is power-limited under the GEMM)
__device__ __forceinline__ bool mbar_try_hint(uint64_t* b, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(b)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
// Whole-warp wait for a long phase (the epilogue waiting for a tile's accumulator): lane 0 sleeps on the barrier, the
// other 31 lanes sit in the shuffle.
template <typename S>
__device__ __forceinline__ bool mbar_wait_warp(S& s, uint64_t* b, uint32_t parity, int lane) {
  uint32_t ok = 1;
  if (lane == 0 && !mbar_try(b, parity)) {
    const unsigned long long t0 = globaltimer_ns();
    while (!mbar_try_hint(b, parity, 20000u)) {
      if (s.abort) { ok = 0; break; }
      if (globaltimer_ns() - t0 > kWaitNs) { s.abort = 1; ok = 0; break; }
    }
  }
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(smem_dst)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B operand tile: 128-byte rows, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc(const void* p) {
  uint64_t d = (uint64_t)((s32(p) & 0x3ffff) >> 4);      // start address, 16-byte units
  d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
  return d;
}
// kind::f16: D=f32, A=B=bf16, both K-major, N=256, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(uint32_t lo_f32, uint32_t hi_f32) {
  __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(lo_f32), __uint_as_float(hi_f32));
  return *reinterpret_cast<uint32_t*>(&v);
}

// bf16 output.  Each epilogue warp owns 32 accumulator rows; per 64-column chunk it converts its rows to
// bf16, lays them into a 4 KiB shared-memory box in the SWIZZLE_128B pattern (16-byte piece j of row r sits
// at piece j ^ (r % 8): conflict-free for the row-per-thread writes) and one lane issues a TMA tensor store.
// The store leaves the SM as whole 128-byte lines, which is what makes the epilogue usable over NVLink
// (`direct` mode: per-thread 16-byte row stores reach a peer as 32 separate small packets per instruction
// -- measured 167 GB/s) and keeps 4x fewer L2 write transactions locally.  Two boxes per warp ping-pong.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(s32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// NBUF staging boxes per warp (2: ping-pong; 1: the wide kernel, whose eight epilogue warps leave 4 KiB each), 64-column
// chunks [c_begin, c_end) of the 256-column accumulator.
struct NoRelease { __device__ __forceinline__ void operator()() const {} };
// `drained()` runs once the last chunk of the range sits in registers (tcgen05.wait::ld passed): the accumulator can be
// handed back to the MMA issuer before the conversion, the staging-box wait and the store of that chunk.
template <int NBUF = 2, typename F = NoRelease>
__device__ __forceinline__ void epilogue_rows_tma(uint8_t (*stage)[kStageC], const CUtensorMap* mc, uint32_t taddr, int col0, int row0, int lane,
                                                  uint32_t dense_n = 0, const float* addend_row = nullptr, uint32_t n_addends = 0,
                                                  uint64_t addend_stride = 0, int c_begin = 0, int c_end = BN / 64, F drained = F()) {
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const uint32_t buf = s32(stage[c % NBUF]);
    uint32_t r[64];
    tmem_ld32(taddr + c * 64, r);                     // in flight while the box is being waited for
    tmem_ld32(taddr + c * 64 + 32, r + 32);
    if (NBUF == 1) {
      tmem_ld_wait();
      if (c == c_end - 1) drained();
      if (lane == 0) bulk_wait_read0();               // the store last issued from this box has read it
      __syncwarp();
    } else {
      if (lane == 0) bulk_wait_read1();               // the store issued from this box two chunks ago has read it
      __syncwarp();
      tmem_ld_wait();
      if (c == c_end - 1) drained();
    }
    for (uint32_t i = 0; i < n_addends; ++i) {         // stream-K: the other clusters' shares of K for this row (fp32, 256 B per chunk)
      const float4* p = reinterpret_cast<const float4*>(addend_row + (uint64_t)i * addend_stride) + (uint64_t)c * 16 * BM;   // [chunk*8 + j][row]
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 v = __ldcg(p + (uint64_t)j * BM);
        r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + v.x);
        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
      }
    }
    const uint32_t rowp = buf + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_v4(rowp + ((uint32_t)(j ^ (lane & 7)) << 4), pack_bf16(r[8 * j], r[8 * j + 1]), pack_bf16(r[8 * j + 2], r[8 * j + 3]),
                   pack_bf16(r[8 * j + 4], r[8 * j + 5]), pack_bf16(r[8 * j + 6], r[8 * j + 7]));
    fence_async_smem();                                // my generic-proxy writes -> visible to the async proxy
    __syncwarp();
    if (lane == 0) {
      if (dense_n) tma_store_2d(mc, stage[c % NBUF], 0, (int)(((uint32_t)row0 / 32u * (dense_n / 64u) + (uint32_t)(col0 + c * 64) / 64u) * 32u));
      else tma_store_2d(mc, stage[c % NBUF], col0 + c * 64, row0);
      bulk_commit();
    }
  }
}

// fp8 output (out_fp8 = 1).  One tcgen05.ld 32x32b.x32 hands each epilogue thread 32 consecutive columns
// of one row -- exactly one MX block -- so the block scale is a register-only reduction: amax over the
// thread's 32 fp32 accumulators, UE8M0 exponent e (smallest power of two with amax / 2^e <= 448, the same
// rule as the pack kernel), 32 x e4m3 = 32 bytes stored, one scale byte.  A 128-row panel becomes one
// self-contained record   [128 x N bytes fp8, row-major][128 x N/32 scale bytes, row-major]
// of 128*N*33/32 bytes, which is what the panel's RDMA write carries: half the wire bytes of bf16.
__device__ __forceinline__ uint64_t panel_record_bytes(uint32_t N) { return (uint64_t)BM * N + (uint64_t)BM * (N / 32); }

__device__ __forceinline__ uint32_t quantize_block(const uint32_t* r, uint32_t* q /* 8 words = 32 x e4m3 */) {
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(__uint_as_float(r[i])));
  const uint32_t vb = __float_as_uint(amax * (1.0f / 448.0f));
  int e = (int)((vb >> 23) & 0xff) - 127 + ((vb & 0x7fffffu) ? 1 : 0);
  e = max(-127, min(127, e));
  const float inv = __uint_as_float((uint32_t)(127 - e) << 23);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(__uint_as_float(r[4 * i]) * inv, __uint_as_float(r[4 * i + 1]) * inv), __NV_SATFINITE, __NV_E4M3);
    uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(__uint_as_float(r[4 * i + 2]) * inv, __uint_as_float(r[4 * i + 3]) * inv), __NV_SATFINITE, __NV_E4M3);
    q[i] = lo | (hi << 16);
  }
  return (uint32_t)(e + 127);
}

// fp8 epilogue through the same staged TMA store as bf16.  A staging box is 32 rows x 128 BYTES = 128 fp8 columns: four
// MX blocks per row, quantised one tcgen05.ld at a time and laid down in the SWIZZLE_128B pattern; one lane stores
// the box through a 3-D tensor map (column, row in panel, panel) over the panel records, which also clips columns
// beyond N and panels beyond M.  The four scale bytes of the box go out as one 32-bit store per row.
// Boxes [b_begin, b_end) of the 256-column accumulator; `drained()` as in epilogue_rows_tma; stream-K addends as there.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(s32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
template <int NBUF = 2, typename F = NoRelease>
__device__ __forceinline__ void epilogue_rows_fp8_tma(uint8_t (*stage)[kStageC], const CUtensorMap* mq, uint32_t taddr, uint32_t col0, uint32_t m_blk,
                                                      uint32_t q, int lane, const GemmArgs& g, int b_begin = 0, int b_end = BN / 128, F drained = F(),
                                                      const float* addend_row = nullptr, uint32_t n_addends = 0, uint64_t addend_stride = 0) {
  const uint32_t n_sc = g.N / 32;
  uint8_t* srow = reinterpret_cast<uint8_t*>(g.c) + (uint64_t)m_blk * panel_record_bytes(g.N) + (size_t)BM * g.N + (size_t)(q * 32 + lane) * n_sc;
  const bool panel_ok = m_blk * BM < g.M;              // a pair / wide tile can reach past the last panel: nothing of it may be written
#pragma unroll 1
  for (int b = b_begin; b < b_end; ++b) {
    const uint32_t rowp = s32(stage[b % NBUF]) + lane * 128;
    uint32_t sc = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(taddr + b * 128 + c * 32, r);
      tmem_ld_wait();
      if (c == 3 && b == b_end - 1) drained();
      if (c == 0) {
        if (lane == 0) { if (NBUF == 2) bulk_wait_read1(); else bulk_wait_read0(); }   // the store last issued from this box has read it
        __syncwarp();
      }
      for (uint32_t i = 0; i < n_addends; ++i) {
        const float4* p = reinterpret_cast<const float4*>(addend_row + (uint64_t)i * addend_stride) + (uint64_t)(b * 4 + c) * 8 * BM;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = __ldcg(p + (uint64_t)j * BM);
          r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + v.x);
          r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
          r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
          r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
        }
      }
      uint32_t w[8];
      sc |= quantize_block(r, w) << (8 * c);
      st_shared_v4(rowp + ((uint32_t)((2 * c) ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
      st_shared_v4(rowp + ((uint32_t)((2 * c + 1) ^ (lane & 7)) << 4), w[4], w[5], w[6], w[7]);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(mq, stage[b % NBUF], (int)(col0 + b * 128), (int)(q * 32), (int)m_blk);
      bulk_commit();
    }
    const uint32_t sc_col = (col0 + (uint32_t)b * 128) / 32;
    if (panel_ok) {
      if (sc_col + 4 <= n_sc && (n_sc & 3) == 0) *reinterpret_cast<uint32_t*>(srow + sc_col) = sc;
      else
        for (uint32_t c = 0; c < 4; ++c)
          if (sc_col + c < n_sc) srow[sc_col + c] = (uint8_t)(sc >> (8 * c));
    }
  }
}
__device__ __forceinline__ void bulk_wait_keep2() { asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); }

// Panel accounting, run by ONE thread after all 128 rows of a tile are written: the CTA that completes a
// 128-row panel's last tile posts the panel (or, in direct mode, its zero-length announcement).
__device__ __forceinline__ void panel_tile_done(const GemmArgs& g, uint32_t m_blk, uint32_t n_blks, bool sys) {
  if (m_blk * BM >= g.M) return;    // a pair / wide tile can reach past the last panel
  fence_gpu();
  unsigned int old = atomicAdd(&g.counters[m_blk], 1u);
  if (old + 1 != n_blks) return;
  fence_scope(sys || g.direct);   // cumulative: covers the other CTAs' tiles of this panel (system scope when they went to a peer)
  const uint32_t rows = min((uint32_t)BM, g.M - m_blk * BM);                  // the last panel may be short (bf16 rows; an fp8 record is always whole)
  const uint64_t stride = g.out_fp8 ? panel_record_bytes(g.N) : (uint64_t)BM * g.N * 2, off = (uint64_t)m_blk * stride;
  const uint64_t panel_bytes = g.direct ? 0 : (g.out_fp8 ? stride : (uint64_t)rows * g.N * 2);
  unsigned long long idx = sq_reserve(g.qp, 1, g.timeout_ns);
  const bool sig = g.signal_every <= 1 || ((idx + 1) % g.signal_every == 0);
  if (idx != ~0ull) {
    write_rdma_wqe(g.qp, idx, g.with_imm ? OP_RDMA_WRITE_IMM : OP_RDMA_WRITE, g.c_va + off, g.lkey, g.remote_va + off, g.rkey,
                   (uint32_t)panel_bytes, sig ? CTRL_CQ_UPDATE : 0, m_blk);
    if (sq_submit(g.qp, idx, 1, g.timeout_ns, true) == WAIT_OK) {
      atomicMax(&g.acc[0], idx + 1);
      atomicAdd(&g.acc[1], 1ull);
      atomicMax(&g.acc[2], ~globaltimer_ns());
    } else g.out[0] = (unsigned long long)(long long)WAIT_TIMEOUT;
  } else g.out[0] = (unsigned long long)(long long)WAIT_TIMEOUT;
  g.counters[m_blk] = 0;
}
__device__ __forceinline__ void bulk_wait_keep4() { asm volatile("cp.async.bulk.wait_group 4;" ::: "memory"); }

// ==================================================================== cta_group::2 variant
// Two CTAs of one cluster (a TPC pair) compute one 256x256 tile: each owns 128 rows of A and of the
// accumulator (its own TMEM) and loads only HALF of the B tile; the 5th-gen tensor core of each SM
// reads the other half from its partner's shared memory.  Per 256x256x64 step a pair moves
// 2 x (16 + 16) KiB instead of 2 x (16 + 32) KiB through L2 -- the 1-CTA kernel already sat at 48 % L2
// throughput at 4096^3 and lost 21 % to cuBLAS at 8192^3 -- and the 32 KiB stages allow a 6-deep ring.
//   * TMA loads carry .cta_group::2 and complete on the LEADER's full barrier (both CTAs' bytes; only the
//     leader arrives on it, the partner's loads just deliver their transaction bytes);
//   * only the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256);
//   * tcgen05.commit multicasts to the same barrier in both CTAs (stage release, accumulator ready);
//   * the partner's epilogue warps hand TMEM back with a remote mbarrier arrive on the leader.
constexpr int STAGES2 = 6;
constexpr int BH_STAGE = (BN / 2) * BK * 2;     // half of the B tile: 16 KiB
struct alignas(1024) Smem2 {
  uint8_t a[STAGES2][A_STAGE];
  uint8_t b[STAGES2][BH_STAGE];
  uint8_t stage_c[4][2][kStageC];
  alignas(8) uint64_t full[STAGES2], empty[STAGES2], tfull[2], tempty[2];
  uint32_t tmem_base;
  volatile int abort;
};
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta address in this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(s32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// Bounded wait of one thread (producer, MMA issuer); false = timed out or another role aborted.  The retry sleeps in the
// hardware (suspend-time hint) instead of spinning: ncu counted 4.4x the library's instructions for the spinning version.
template <typename S>
__device__ __forceinline__ bool mbar_wait_t(S& s, uint64_t* b, uint32_t parity) {
  if (mbar_try(b, parity)) return true;
  unsigned long long t0 = globaltimer_ns();
  unsigned n = 0;
  while (!mbar_try_hint(b, parity, 10000u)) {
    if ((++n & 7) == 0) {
      if (s.abort) return false;
      if (globaltimer_ns() - t0 > kWaitNs) { s.abort = 1; return false; }
    }
  }
  return true;
}

// ------------------------------------------------------------------ the tile kernel: 1 CTA (128x256) or a CTA pair (256x256)
// One body for both shapes (CTAS = 1: cta_group::1, 4 x 48 KiB stages; CTAS = 2: cta_group::2, 6 x 32 KiB stages, see the
// notes above Smem2).  What differs is spelled with `if constexpr`: which CTA arrives / issues, the .cta_group of the
// TMA / MMA / commit / alloc instructions, and local vs remote mbarrier arrives.  Shapes need not be tile multiples:
// the tile grid is a ceiling division, TMA zero-fills what it reads out of bounds (rows of A / B and the K tail) and
// clips what it stores out of bounds; only the panel accounting knows about the short last panel.
template <int CTAS> struct SmemSel;
template <> struct SmemSel<1> { using type = Smem; static constexpr int kStages = STAGES; static constexpr int kBStage = B_STAGE; };
template <> struct SmemSel<2> { using type = Smem2; static constexpr int kStages = STAGES2; static constexpr int kBStage = BH_STAGE; };

// last CTA out: flush NOP (its CQE = everything before it was delivered), result words, scratch reset
__device__ __forceinline__ void finish_kernel(const GemmArgs& g, bool aborted, unsigned long long t_start, uint32_t variant) {
  const uint32_t panels = (g.M + BM - 1) / BM;
  if (aborted) g.out[0] = (unsigned long long)(long long)WAIT_TIMEOUT;
  fence_gpu();
  unsigned int old = atomicAdd(&g.counters[panels], 1u);
  if (old + 1 != gridDim.x) return;
  const unsigned long long t_compute_end = globaltimer_ns();
  fence_gpu();
  unsigned long long posted = ld_u64_volatile(&g.acc[1]);
  if (g.qp != nullptr) {
    int rc = WAIT_TIMEOUT;
    unsigned long long fidx = sq_reserve(g.qp, 1, g.timeout_ns);
    if (fidx != ~0ull) {
      uint8_t* slot = g.qp->sq + ((fidx & ((1ull << g.qp->sq_log) - 1)) << 6);
      st_v4(slot + 0, ctrl_word0(OP_NOP, (uint16_t)fidx), ctrl_word1(g.qp->qpn, 1), (uint32_t)CTRL_CQ_UPDATE << 24, 0u);
      st_v4(slot + 16, 0u, 0u, 0u, 0u);
      st_v4(slot + 32, 0u, 0u, 0u, 0u);
      st_v4(slot + 48, 0u, 0u, 0u, 0u);
      if (sq_submit(g.qp, fidx, 1, g.timeout_ns, true) == WAIT_OK) rc = g.post_only ? WAIT_OK : sq_wait(g.qp, fidx, g.timeout_ns);
    }
    if (posted != panels && rc == WAIT_OK) rc = WAIT_TIMEOUT;
    if (rc != WAIT_OK) g.out[0] = (unsigned long long)(long long)rc;
  }
  g.out[1] = t_start; g.out[2] = globaltimer_ns(); g.out[3] = posted;
  g.out[4] = ~ld_u64_volatile(&g.acc[2]); g.out[5] = t_compute_end; g.out[6] = variant;
  g.counters[panels] = 0;
  g.acc[0] = 0; g.acc[1] = 0; g.acc[2] = 0;
}

template <int CTAS>
__device__ __forceinline__ void gemm_tile_body(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const CUtensorMap& tmap_c, const GemmArgs& g) {
  using S = typename SmemSel<CTAS>::type;
  constexpr int kStages = SmemSel<CTAS>::kStages, kBStage = SmemSel<CTAS>::kBStage;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  S& s = *reinterpret_cast<S*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_rank() : 0u;
  const bool leader = rank == 0;
  const unsigned long long t_start = globaltimer_ns();
  const uint32_t m_tiles = (g.M + CTAS * BM - 1) / (CTAS * BM), n_blks = (g.N + BN - 1) / BN, k_blks = (g.K + BK - 1) / BK;
  const uint32_t n_tiles = m_tiles * n_blks, n_units = gridDim.x / CTAS, unit = blockIdx.x / CTAS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s.tfull[i], 1); mbar_init(&s.tempty[i], 4 * CTAS); }
    s.abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1) {   // whole warp (.sync.aligned); same warp id in both CTAs of a pair
    if constexpr (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s.tmem_base)), "r"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s.tmem_base)), "r"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();      // partner's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (every CTA: its own 128 rows of A, its own 1/CTAS of the B tile)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t tile = unit; tile < n_tiles && !s.abort; tile += n_units) {
        uint32_t mt, n_blk;
        tile_coords(tile, m_tiles, n_blks, g.group_m, &mt, &n_blk);
        const int a_row = (int)((mt * CTAS + rank) * BM), b_row = (int)(n_blk * BN + rank * (BN / CTAS));
        for (uint32_t kb = 0; kb < k_blks; ++kb) {
          if (!mbar_wait_t(s, &s.empty[stage], phase ^ 1)) goto producer_done;
          if constexpr (CTAS == 2) {
            // Only the leader arrives (expecting BOTH CTAs' bytes); the partner's TMA completes its bytes on the
            // leader's barrier by itself.  A remote release-arrive from the partner here cost a cluster-scope
            // memory barrier per k-block and halved the kernel (ncu: tensor pipe 35 %).
            const uint32_t lbar = mapa(s32(&s.full[stage]), 0);
            if (leader) mbar_expect_tx(&s.full[stage], 2 * (A_STAGE + kBStage));
            tma_load_2d_2sm(s.a[stage], &tmap_a, lbar, (int)(kb * BK), a_row);
            tma_load_2d_2sm(s.b[stage], &tmap_b, lbar, (int)(kb * BK), b_row);
          } else {
            mbar_expect_tx(&s.full[stage], A_STAGE + kBStage);
            tma_load_2d(s.a[stage], &tmap_a, &s.full[stage], (int)(kb * BK), a_row);
            tma_load_2d(s.b[stage], &tmap_b, &s.full[stage], (int)(kb * BK), b_row);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  producer_done:
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; of the leader CTA in a pair)
    if (leader && lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t tile = unit; tile < n_tiles && !s.abort; tile += n_units) {
        if (!mbar_wait_t(s, &s.tempty[acc], acc_phase ^ 1)) goto mma_done;   // every epilogue warp drained this buffer
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (uint32_t kb = 0; kb < k_blks; ++kb) {
          if (!mbar_wait_t(s, &s.full[stage], phase)) goto mma_done;          // A and B (both CTAs' halves) landed
          tc_fence_after();
          const uint64_t da = smem_desc(s.a[stage]), db = smem_desc(s.b[stage]);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {                              // +32 bytes along K = +2 in the address field
            if constexpr (CTAS == 2) umma_f16_2sm(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc2, (kb | (uint32_t)k) != 0);
            else umma_f16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc, (kb | (uint32_t)k) != 0);
          }
          if constexpr (CTAS == 2) tc_commit_2sm(&s.empty[stage]); else tc_commit(&s.empty[stage]);   // frees the stage when these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (CTAS == 2) tc_commit_2sm(&s.tfull[acc]); else tc_commit(&s.tfull[acc]);         // accumulator complete (in both CTAs)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  mma_done:
    __syncwarp();
  } else {
    // ===================== epilogue (4 warps <-> 4 TMEM lane quarters; each CTA its own 128 rows)
    const uint32_t q = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    const bool sys = g.qp != nullptr && poster_sys(g.qp);
    const bool defer = !g.plain_stores;
    uint32_t pending = ~0u;
    for (uint32_t tile = unit; tile < n_tiles; tile += n_units) {
      uint32_t mt, n_blk;
      tile_coords(tile, m_tiles, n_blks, g.group_m, &mt, &n_blk);
      const uint32_t m_blk = mt * CTAS + rank;                                 // 128-row panel index
      if (!mbar_wait_warp(s, &s.tfull[acc], acc_phase, lane)) break;
      tc_fence_after();
      const uint32_t row_in_panel = q * 32 + lane;
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * BN;
      if (g.out_fp8) {
        epilogue_rows_fp8_tma(s.stage_c[q], &tmap_c, taddr, n_blk * BN, m_blk, q, lane, g);
      } else if (!g.plain_stores) {
        epilogue_rows_tma(s.stage_c[q], &tmap_c, taddr, (int)(n_blk * BN), (int)(m_blk * BM + q * 32), lane, g.dense_probe ? g.N : 0u);
      } else {                                                                 // A/B switch, tile-multiple shapes only (host checks)
        __nv_bfloat16* crow = g.c + (size_t)(m_blk * BM + row_in_panel) * g.N + (size_t)n_blk * BN;
#pragma unroll 2
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          uint4* dst = reinterpret_cast<uint4*>(crow + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(pack_bf16(r[8 * j], r[8 * j + 1]), pack_bf16(r[8 * j + 2], r[8 * j + 3]),
                                pack_bf16(r[8 * j + 4], r[8 * j + 5]), pack_bf16(r[8 * j + 6], r[8 * j + 7]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                                         // 4 x CTAS arrivals hand the buffer back (on the LEADER's barrier)
        if (leader) mbar_arrive(&s.tempty[acc]);
        else mbar_arrive_remote(mapa(s32(&s.tempty[acc]), 0));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (g.qp != nullptr) {
        // ---- panel accounting.  With staged TMA stores it runs ONE TILE LATE: tile i's store groups (four bf16 boxes
        // or two fp8 boxes per warp) stay in flight -- a peer acknowledges them microseconds later -- while tile i-1
        // is accounted for.
        if (defer) {
          if (pending != ~0u) {
            if (lane == 0) { if (g.out_fp8) bulk_wait_keep2(); else bulk_wait_keep4(); }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (threadIdx.x == 64) panel_tile_done(g, pending, n_blks, sys);
          }
          pending = m_blk;
        } else {
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // all 128 rows of this tile are stored
          if (threadIdx.x == 64) panel_tile_done(g, m_blk, n_blks, sys);
        }
      }
    }
    if (lane == 0) bulk_wait_all();                                             // staged rows have left shared memory (and are written)
    if (pending != ~0u && !s.abort) {
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (threadIdx.x == 64) panel_tile_done(g, pending, n_blks, sys);
    }
  }

  // ===================== teardown
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();      // nobody frees TMEM while the partner's tensor core may still write it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
  if (threadIdx.x == 64) finish_kernel(g, s.abort != 0, t_start, (uint32_t)CTAS);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_send_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ GemmArgs g) {
  gemm_tile_body<1>(tmap_a, tmap_b, tmap_c, g);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_send2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ GemmArgs g) {
  gemm_tile_body<2>(tmap_a, tmap_b, tmap_c, g);
}

// ------------------------------------------------------------------ wide variant: 256x256 per CTA, 512x256 per CTA pair
// What ncu showed against cuBLAS at 8192^3 (profiles/ncu/gemm_vs_cublas_*): the pair kernel above needs 1.4 % more
// cycles than nvjet_tst_256x256_64x4_2x1_2cta -- and runs them at 1.43 GHz where the library holds 1.50 GHz: the box
// is power-limited, and the pair kernel moves 1.7x the L2 -> SM bytes (252 M vs 149 M sectors) for the same FLOPs.
// The library's shape is 256 rows of A per CTA: per 64-deep k-block a CTA loads 32 KiB of A and its 16 KiB half of
// B for TWO M=256 pair-MMAs (rows 0-127 and 128-255 of each CTA, accumulators in TMEM columns 0-255 and 256-511),
// i.e. 48 KiB per 8.4 MFLOP instead of 32 KiB per 4.2: a quarter less operand traffic, half the barrier and TMA
// operations per FLOP.  TMEM is then single-buffered (512 columns = one 256x256 fp32 tile), so the epilogue hands the
// two halves back separately: the next tile's MMAs on columns 0-255 start while columns 256-511 are still being
// drained, and the TMA producer keeps prefetching the next tile's stages throughout (the kernel is persistent,
// unlike the library's one-tile CTAs).
// Epilogue.  With four epilogue warps and the MMA loop waiting for half 1 right after its first four MMAs on half 0, the
// whole drain of a tile (~4 k cycles) was exposed: 3 % of a K = 8192 tile, 6 % at K = 4096 -- the gap to the library.
// Now (a) eight epilogue warps: the two warps of a TMEM lane quarter take 128 columns each, so a half drains in half the
// time; (b) the MMA issuer runs half 0 up to STAGES3 - 1 k-blocks ahead while half 1 is still being drained (the
// operand stages are simply released later, after half 1 has consumed them too).  Exposed: one half-drain by 8 warps.
constexpr int STAGES3 = 4;
constexpr int A3_STAGE = 2 * A_STAGE;           // 256 rows x 64: 32 KiB
// Timeline of cluster 0's leader CTA, %globaltimer ns (rn_gemm_timeline reads it): [0] kernel entry, [1] setup done,
// [2] first operands landed, [3] last MMA issued, [4] first tile's accumulator complete, [5] epilogue warp 2 done with its
// last tile, [6] CTA exit, [7] clock64 cycles of the issuer loop.  A handful of stores per launch.
__device__ unsigned long long g_timeline[16];   // [8], [9]: cycles from the first tile's accumulator-ready to warp 2 having stored its part of half 0 / half 1
constexpr int kThreads3 = 320, kEpiThreads3 = 256, kEpiWarps3 = 8;
struct alignas(1024) Smem3 {
  uint8_t a[STAGES3][A3_STAGE];
  uint8_t b[STAGES3][BH_STAGE];
  uint8_t stage_c[kEpiWarps3][1][kStageC];
  alignas(8) uint64_t full[STAGES3], empty[STAGES3], tfull, tempty[2];
  uint32_t tmem_base;
  volatile int abort;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads3, 1)
gemm_send3_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem3& s = *reinterpret_cast<Smem3*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const unsigned long long t_start = globaltimer_ns();
  const uint32_t mq_blks = (g.M + 4 * BM - 1) / (4 * BM), n_blks = (g.N + BN - 1) / BN, k_blks = (g.K + BK - 1) / BK;
  const uint32_t n_tiles = mq_blks * n_blks, n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  // Schedule.  Whole tiles go round robin (cluster c takes tiles c, c + n_clusters, ...): clusters that run at the
  // same time sit on neighbouring tiles of the grouped rasterisation, which keeps the operands they share in L2 (an
  // even split of the whole unit sequence was tried first and lost 25 %: it shifts every cluster's start by a
  // fraction of a tile, so concurrent clusters end up waves apart).  What round robin leaves is the last, partial
  // wave of R = n_tiles % n_clusters tiles, during which n_clusters - R cluster pairs idle: 13.5 % of the run at
  // 4096^3.  Stream-K (g.ws != nullptr) splits exactly that wave: its R * k_blks (tile, k-block) units are dealt out
  // evenly to ALL clusters.  A tail tile is then computed by a run of neighbouring clusters; the one that has its
  // first k-blocks owns it and folds the others' fp32 partials (parked in the workspace) into its epilogue.  A
  // cluster's share is shorter than a tile, so it holds at most the tail end of one tile (computed first, parked)
  // and the head of the next (computed last, owned): nobody waits on work that has not started.
  const uint32_t W = n_tiles / n_clusters, R = n_tiles % n_clusters;
  const bool streamk = g.ws != nullptr && R != 0;
  const uint64_t tail_units = (uint64_t)R * k_blks;
  // clusters that take part in the tail: all of them, unless that would leave shares of under 4 k-blocks (each
  // share costs its owner one more partial to fold in); never fewer than R (then nothing is split at all)
  uint32_t n_tail = n_clusters;
  if (tail_units / 4 < n_tail) n_tail = (uint32_t)(tail_units / 4);
  if (n_tail < R) n_tail = R;
  const bool in_tail = cluster_id < n_tail;
  const uint64_t t_begin = in_tail ? tail_units * cluster_id / n_tail : 0, t_end = in_tail ? tail_units * (cluster_id + 1) / n_tail : 0;
  struct Seg { uint32_t tile, kb0, kb1, j; };
  struct Cursor { uint32_t wave; uint64_t ut; };
  auto next_seg = [&](Cursor& cur, Seg& sg) -> bool {
    if (cur.wave < W) {
      sg.tile = cur.wave * n_clusters + cluster_id; sg.kb0 = 0; sg.kb1 = k_blks; sg.j = ~0u;
      ++cur.wave;
      return true;
    }
    if (!streamk) {
      if (cur.wave != W || cluster_id >= R) return false;
      sg.tile = W * n_clusters + cluster_id; sg.kb0 = 0; sg.kb1 = k_blks; sg.j = ~0u;
      ++cur.wave;
      return true;
    }
    if (cur.ut >= t_end) return false;
    sg.j = (uint32_t)(cur.ut / k_blks);
    sg.tile = W * n_clusters + sg.j;
    sg.kb0 = (uint32_t)(cur.ut % k_blks);
    const uint64_t left = t_end - cur.ut;
    sg.kb1 = (uint32_t)((uint64_t)sg.kb0 + left < k_blks ? sg.kb0 + left : k_blks);
    cur.ut += sg.kb1 - sg.kb0;
    return true;
  };
  // cluster that holds tail unit x (largest c with tail_units * c / n_tail <= x)
  auto cluster_of = [&](uint64_t x) -> uint32_t {
    uint32_t c = (uint32_t)(x * n_tail / tail_units) + 1;
    if (c > n_tail - 1) c = n_tail - 1;
    while (tail_units * c / n_tail > x) --c;
    return c;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES3; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    mbar_init(&s.tfull, 1);
    mbar_init(&s.tempty[0], 2 * kEpiWarps3); mbar_init(&s.tempty[1], 2 * kEpiWarps3);
    s.abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s.tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;
  const bool tl = cluster_id == 0 && leader;
  if (tl && threadIdx.x == 0) { g_timeline[0] = t_start; g_timeline[1] = globaltimer_ns(); }

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own 256 rows of A, own half of B)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      Cursor cur = {0, t_begin};
      Seg sg;
      while (!s.abort && next_seg(cur, sg)) {
        uint32_t mq, n_blk;
        tile_coords(sg.tile, mq_blks, n_blks, g.group_m, &mq, &n_blk);
        for (uint32_t kb = sg.kb0; kb < sg.kb1; ++kb) {
          if (!mbar_wait_t(s, &s.empty[stage], phase ^ 1)) goto producer3_done;
          const uint32_t lbar = mapa(s32(&s.full[stage]), 0);
          if (leader) mbar_expect_tx(&s.full[stage], 2 * (A3_STAGE + BH_STAGE));
          tma_load_2d_2sm(s.a[stage], &tmap_a, lbar, (int)(kb * BK), (int)(mq * 4 * BM + rank * 2 * BM));
          tma_load_2d_2sm(s.b[stage], &tmap_b, lbar, (int)(kb * BK), (int)(n_blk * BN + rank * (BN / 2)));
          if (++stage == STAGES3) { stage = 0; phase ^= 1; }
        }
      }
    }
  producer3_done:
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only): two M=256 pair-MMAs per k16, one per accumulator half
    if (leader && lane == 0) {
      uint32_t stage = 0, phase = 0, tphase = 0;
      Cursor cur = {0, t_begin};
      Seg sg;
      auto issue_half = [&](uint32_t st, int h, bool first) {
        const uint64_t da = smem_desc(s.a[st] + h * A_STAGE), db = smem_desc(s.b[st]);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_f16_2sm(tmem_base + (uint32_t)h * BN, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc2, !first || k != 0);
      };
      // where the issuer's time goes (cluster 0 reports it in out[7]): cycles spent waiting for operands / for TMEM
      long long w_full = 0, w_tmem = 0;
      bool landed = false;
      const long long t_loop = clock64();
      while (!s.abort && next_seg(cur, sg)) {
        const uint32_t nkb = sg.kb1 - sg.kb0;
        const uint32_t ahead = nkb < (uint32_t)(STAGES3 - 1) ? nkb : (uint32_t)(STAGES3 - 1);
        // head of the tile: half 0 as soon as both CTAs' epilogues have drained it, `ahead` k-blocks deep ...
        long long t0 = clock64();
        if (!mbar_wait_t(s, &s.tempty[0], tphase ^ 1)) goto mma3_done;
        w_tmem += clock64() - t0;
        tc_fence_after();
        uint32_t st = stage, ph = phase;
        for (uint32_t i = 0; i < ahead; ++i) {
          t0 = clock64();
          if (!mbar_wait_t(s, &s.full[st], ph)) goto mma3_done;
          w_full += clock64() - t0;
          if (tl && !landed) { g_timeline[2] = globaltimer_ns(); landed = true; }
          tc_fence_after();
          issue_half(st, 0, i == 0);
          if (++st == STAGES3) { st = 0; ph ^= 1; }
        }
        // ... then half 1 catches up on the same stages and releases them
        t0 = clock64();
        if (!mbar_wait_t(s, &s.tempty[1], tphase ^ 1)) goto mma3_done;
        w_tmem += clock64() - t0;
        tc_fence_after();
        for (uint32_t i = 0; i < ahead; ++i) {
          issue_half(stage, 1, i == 0);
          tc_commit_2sm(&s.empty[stage]);
          if (++stage == STAGES3) { stage = 0; phase ^= 1; }
        }
        for (uint32_t kb = sg.kb0 + ahead; kb < sg.kb1; ++kb) {
          t0 = clock64();
          if (!mbar_wait_t(s, &s.full[stage], phase)) goto mma3_done;
          w_full += clock64() - t0;
          tc_fence_after();
          issue_half(stage, 0, false);
          issue_half(stage, 1, false);
          tc_commit_2sm(&s.empty[stage]);
          if (++stage == STAGES3) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(&s.tfull);                                               // both halves complete, in BOTH CTAs
        tphase ^= 1;
      }
      if (cluster_id == 0) {
        const unsigned long long tot = (unsigned long long)(clock64() - t_loop);
        g_timeline[3] = globaltimer_ns(); g_timeline[7] = tot;
        g.out[7] = (((unsigned long long)w_full >> 4) & 0x1fffffull) | ((((unsigned long long)w_tmem >> 4) & 0x1fffffull) << 21) | ((tot >> 4) << 42);
      }
    }
  mma3_done:
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs: their own 256 rows = two 128-row panels of the tile)
    // warps 2..9: TMEM lane quarter q = warp % 4 (hardware rule), column part cp: 128 of a half's 256 columns
    const uint32_t q = warp & 3, cp = (uint32_t)(warp - 2) >> 2;
    uint32_t tphase = 0;
    Cursor cur = {0, t_begin};
    Seg sg;
    const bool sys = g.qp != nullptr && poster_sys(g.qp);
    uint32_t pending0 = ~0u, pending1 = ~0u;
    bool first_tile = true, first_tile_epi = true;
    constexpr uint64_t kWsPerCta = 2ull * BM * BN;                              // floats: two 128 x 256 halves
    while (next_seg(cur, sg)) {
      uint32_t mq, n_blk;
      tile_coords(sg.tile, mq_blks, n_blks, g.group_m, &mq, &n_blk);
      if (!mbar_wait_warp(s, &s.tfull, tphase, lane)) break;
      if (tl && threadIdx.x == 64 && first_tile) { g_timeline[4] = globaltimer_ns(); first_tile = false; }
      const long long t_wake = clock64();
      tphase ^= 1;
      tc_fence_after();
      const bool contributor = sg.kb0 != 0;                                    // not the head of its tile: park the partial, store nothing
      const uint32_t row_in_panel = q * 32 + lane;
      if (contributor) {
        float* wsb = g.ws + ((uint64_t)cluster_id * 2 + rank) * kWsPerCta;     // one slot per contributing cluster
#pragma unroll 1
        for (uint32_t h = 0; h < 2; ++h) {
          const uint32_t taddr = tmem_base + ((q * 32u) << 16) + h * BN;
          // layout private to this kernel, chosen for the accesses: [half][32-column chunk][float4 j][row] -- the 32 rows
          // of a warp are 512 contiguous bytes per store / load instruction
          float4* dst = reinterpret_cast<float4*>(wsb) + (uint64_t)h * (BN / 4) * BM + row_in_panel;
#pragma unroll 2
          for (int c = (int)cp * 4; c < (int)cp * 4 + 4; ++c) {
            uint32_t r[32];
            tmem_ld32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              __stcg(dst + (uint64_t)(c * 8 + j) * BM, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                                  __uint_as_float(r[4 * j + 3])));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&s.tempty[h]);
            else mbar_arrive_remote(mapa(s32(&s.tempty[h]), 0));
          }
        }
        __threadfence();                                                       // my rows of the partial before the flag
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads3) : "memory");
        if (threadIdx.x == 64)
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(g.ws_flags + (uint64_t)cluster_id * 2 + rank), "r"(g.epoch) : "memory");
        continue;
      }
      // head of a split tail tile: the clusters after this one computed the rest of K at the same time
      uint32_t n_contrib = 0;
      if (sg.j != ~0u && sg.kb1 != k_blks) n_contrib = cluster_of((uint64_t)(sg.j + 1) * k_blks - 1) - cluster_id;
      for (uint32_t i = 1; i <= n_contrib && !s.abort; ++i) {
        const unsigned int* flag = g.ws_flags + (uint64_t)(cluster_id + i) * 2 + rank;
        unsigned long long t0 = 0;
        unsigned int v;
        for (unsigned n = 0;; ++n) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          if (v == g.epoch || s.abort) break;
          if ((n & 63) == 63) {
            if (!t0) t0 = globaltimer_ns();
            else if (globaltimer_ns() - t0 > kWaitNs) { s.abort = 1; break; }
          }
        }
      }
      const float* wsa = n_contrib ? g.ws + ((uint64_t)(cluster_id + 1) * 2 + rank) * kWsPerCta : nullptr;
#pragma unroll 1
      for (uint32_t h = 0; h < 2; ++h) {
        const uint32_t m_blk = mq * 4 + rank * 2 + h;                          // 128-row panel index
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + h * BN;
        const float* add_row = wsa ? wsa + ((uint64_t)h * (BN / 4) * BM + row_in_panel) * 4 : nullptr;   // float4 [chunk*8 + j][row], see the contributor
        auto release = [&]() {                                                 // 16 arrivals (8 warps x 2 CTAs) on the LEADER's barrier
          if (tl && warp == 2 && first_tile_epi) g_timeline[10 + h] = (unsigned long long)(clock64() - t_wake);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&s.tempty[h]);
            else mbar_arrive_remote(mapa(s32(&s.tempty[h]), 0));
          }
        };
        if (!g.out_fp8) {
          epilogue_rows_tma<1>(s.stage_c[warp - 2], &tmap_c, taddr, (int)(n_blk * BN), (int)(m_blk * BM + q * 32), lane, 0u, add_row, n_contrib,
                               2 * kWsPerCta, (int)cp * 2, (int)cp * 2 + 2, release);
          if (tl && warp == 2 && first_tile_epi) { g_timeline[8 + h] = (unsigned long long)(clock64() - t_wake); }
        } else {
          epilogue_rows_fp8_tma<1>(s.stage_c[warp - 2], &tmap_c, taddr, n_blk * BN, m_blk, q, lane, g, (int)cp, (int)cp + 1, release, add_row, n_contrib,
                                   2 * kWsPerCta);
        }
      }
      first_tile_epi = false;
      if (g.qp != nullptr) {
        // panel accounting, one tile late for the staged TMA stores (see the tile kernel): a warp issues four bf16 or two
        // fp8 boxes per tile, so "all but the newest 4 / 2 groups complete" = everything of the previous tile has landed
        if (pending0 != ~0u) {
          if (lane == 0) { if (g.out_fp8) bulk_wait_keep2(); else bulk_wait_keep4(); }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads3) : "memory");
          if (threadIdx.x == 64) { panel_tile_done(g, pending0, n_blks, sys); panel_tile_done(g, pending1, n_blks, sys); }
        }
        pending0 = mq * 4 + rank * 2; pending1 = pending0 + 1;
      }
    }
    if (lane == 0) bulk_wait_all();
    if (tl && threadIdx.x == 64) g_timeline[5] = globaltimer_ns();
    if (pending0 != ~0u && !s.abort) {
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads3) : "memory");
      if (threadIdx.x == 64) { panel_tile_done(g, pending0, n_blks, sys); panel_tile_done(g, pending1, n_blks, sys); }
    }
  }

  // ===================== teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
  if (threadIdx.x == 64) {
    if (tl) g_timeline[6] = globaltimer_ns();
    finish_kernel(g, s.abort != 0, t_start, 3u);
  }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// row-major [rows, K] bf16, box = [box_rows, 64], 128-byte swizzle (operand loads, and the C store map with 32-row boxes)
int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return -38;
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {K * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 2000;
}

// fp8 panel records as a 3-D tensor: (column, row in panel, panel) with strides (1, N, record bytes); box = 128 columns x 32
// rows of one panel, 128-byte swizzle -- the fp8 epilogue's staging box
int make_map_records(CUtensorMap* m, const void* base, uint64_t N, uint64_t panels) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return -38;
  cuuint64_t dims[3] = {N, (cuuint64_t)BM, panels};
  cuuint64_t strides[2] = {N, (cuuint64_t)BM * N + (cuuint64_t)BM * (N / 32)};
  cuuint32_t box[3] = {128, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 2000;
}

}  // namespace

RN_API uint64_t rn_gemm_panel_record_bytes(uint32_t N) { return (uint64_t)BM * N + (uint64_t)BM * (N / 32); }
RN_API uint32_t rn_gemm_tile(uint32_t* bm, uint32_t* bn, uint32_t* bk) { *bm = BM; *bn = BN; *bk = BK; return STAGES; }

// counters_dev: >= (M/128 + 1) * 4 + 32 bytes of zeroed device scratch (self-cleaning), out_dev: 64 B mapped pinned.
// Stream-K workspace of the wide kernel: one slot per cluster boundary (at most 73 on 148 SMs) of 2 CTAs x 2 x 128 x 256
// fp32, followed by the flags.  Set per device by the context that owns the memory (hca_host.cu allocates it with the
// HCA: an allocation at launch time would stall behind a resident engine kernel).
namespace { constexpr uint32_t kWsSlots = 80; constexpr uint64_t kWsSlotFloats = 2ull * 2 * BM * BN; }
static float* g_ws[16] = {};
static uint32_t g_epoch = 1;
RN_API int rn_gemm_timeline(unsigned long long* out8 /* [16] */) {
  return (int)cudaMemcpyFromSymbol(out8, g_timeline, sizeof(unsigned long long) * 16);
}
RN_API uint64_t rn_gemm_workspace_bytes() { return kWsSlots * kWsSlotFloats * 4 + kWsSlots * 2 * 4 + 256; }
RN_API void rn_gemm_set_workspace(int dev, uint64_t ptr) { if (dev >= 0 && dev < 16) g_ws[dev] = (float*)ptr; }

// Shapes: any M; N % 8 == 0 and K % 8 == 0 (16-byte rows for TMA); fp8 output additionally N % 32 == 0 (whole MX blocks).
// Tiles that hang over an edge are zero-filled on load and clipped on store by TMA.  The plain-store / probe / direct
// variants are measurement switches and keep the tile-multiple restriction.
RN_API int rn_k_gemm_send(uint64_t stream, int grid, uint64_t a, uint64_t b, uint64_t c, uint32_t M, uint32_t N, uint32_t K,
                          uint64_t qp_dev, uint64_t c_va, uint32_t lkey, uint64_t remote_va, uint32_t rkey,
                          uint32_t signal_every, uint32_t with_imm, uint32_t out_fp8, uint32_t cta_group, uint32_t group_m, uint32_t flags, uint64_t counters_dev,
                          uint64_t out_dev, uint64_t timeout_ms) {
  if (!M || !N || !K || N % 8 || K % 8) return -22;
  if (out_fp8 && N % 32) return -22;
  if ((a | b | c) & 15) return -22;
  const bool aligned = M % BM == 0 && N % BN == 0 && K % BK == 0;
  if ((flags & (kFlagPlainStores | kFlagDenseProbe | kFlagDirect)) && !aligned) return -22;
  if ((flags & (kFlagPlainStores | kFlagDenseProbe)) && out_fp8) return -22;
  const uint32_t panels = (M + BM - 1) / BM, n_blks = (N + BN - 1) / BN;
  CUtensorMap ma, mb, mc;
  int rc = out_fp8 ? make_map_records(&mc, (const void*)c, N, panels)
           : (flags & kFlagDenseProbe) ? make_map(&mc, (const void*)c, (uint64_t)M * N / 64, 64, 32)
                                       : make_map(&mc, (const void*)c, M, N, 32);
  if (rc) return rc;
  GemmArgs g;
  g.ws = nullptr; g.ws_flags = nullptr; g.epoch = 0;
  g.c = (__nv_bfloat16*)c; g.M = M; g.N = N; g.K = K; g.qp = (QpDev*)qp_dev; g.c_va = c_va; g.lkey = lkey; g.rkey = rkey;
  g.remote_va = remote_va; g.signal_every = signal_every ? signal_every : 1; g.with_imm = with_imm & 1u; g.post_only = (with_imm >> 1) & 1u; g.out_fp8 = out_fp8; g.group_m = group_m ? group_m : 1;
  g.direct = (flags & kFlagDirect) ? 1 : 0; g.plain_stores = (flags & kFlagPlainStores) ? 1 : 0; g.dense_probe = (flags & kFlagDenseProbe) ? 1 : 0;
  g.counters = (unsigned int*)counters_dev;
  g.acc = (unsigned long long*)(counters_dev + (((uint64_t)panels + 1) * 4 + 7) / 8 * 8);
  g.out = (unsigned long long*)out_dev;
  g.timeout_ns = (timeout_ms ? timeout_ms : 2000) * 1000000ull;
  unsigned long long* o = (unsigned long long*)out_dev;
  for (int i = 0; i < 8; ++i) o[i] = 0;
  if (grid <= 0) grid = 148;
  // cta_group: 1 = one CTA per 128x256 tile, 2 = CTA pair per 256x256 tile, 3 = wide CTA pair (512x256), 0 = the widest
  // tile that M fills exactly (a tile hanging over the last rows computes zeros: correct, but wasted tensor time).
  // The wide kernel has no plain-store / probe epilogues and no direct mode.
  const bool wide_ok = grid >= 2 && !(flags & (kFlagPlainStores | kFlagDenseProbe | kFlagDirect));
  if ((cta_group == 3 || (cta_group == 0 && M % (4 * BM) == 0)) && wide_ok) {
    rc = make_map(&ma, (const void*)a, M, K, 2 * BM);                 // 256 rows of A per CTA in one box
    if (!rc) rc = make_map(&mb, (const void*)b, N, K, BN / 2);
    if (rc) return rc;
    const uint32_t n_tiles = ((M + 4 * BM - 1) / (4 * BM)) * n_blks;
    grid &= ~1;
    if ((uint32_t)grid > 2 * n_tiles) grid = (int)(2 * n_tiles);
    if (g.group_m > 1) g.group_m = (g.group_m + 1) / 2;              // group_m is given in 256-row units; tiles here are 512 rows
    {
      // stream-K needs every cluster resident at once (a tile's two halves are computed by neighbouring clusters):
      // true for grid <= #SMs with one CTA per SM, which the caller controls; kFlagNoStreamK opts out
      int dev = 0;
      cudaGetDevice(&dev);
      float* ws = (dev >= 0 && dev < 16) ? g_ws[dev] : nullptr;
      // worth it only when whole tiles do not divide evenly among the clusters (round robin is then perfectly balanced)
      const bool want = !(flags & kFlagNoStreamK) && ws && (uint32_t)(grid / 2) <= kWsSlots && n_tiles % (uint32_t)(grid / 2) != 0;
      g.ws = want ? ws : nullptr;
      g.ws_flags = want ? (unsigned int*)(ws + kWsSlots * kWsSlotFloats) : nullptr;
      if (++g_epoch == 0) g_epoch = 1;
      g.epoch = g_epoch;
    }
    const size_t smem = sizeof(Smem3) + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm_send3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    gemm_send3_kernel<<<grid, kThreads3, smem, (cudaStream_t)stream>>>(ma, mb, mc, g);
    return (int)cudaGetLastError();
  }
  rc = make_map(&ma, (const void*)a, M, K, BM);
  if (rc) return rc;
  const bool two = (cta_group == 2 || cta_group == 3 || (cta_group == 0 && M % (2 * BM) == 0)) && grid >= 2;
  if (two) {
    // pairs of CTAs share a 256x256 tile: B map delivers half tiles (128 rows)
    rc = make_map(&mb, (const void*)b, N, K, BN / 2);
    if (rc) return rc;
    const uint32_t n_tiles = ((M + 2 * BM - 1) / (2 * BM)) * n_blks;
    grid &= ~1;
    if ((uint32_t)grid > 2 * n_tiles) grid = (int)(2 * n_tiles);
    const size_t smem = sizeof(Smem2) + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm_send2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    gemm_send2_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma, mb, mc, g);
    return (int)cudaGetLastError();
  }
  rc = make_map(&mb, (const void*)b, N, K, BN);
  if (rc) return rc;
  const uint32_t n_tiles = panels * n_blks;
  if ((uint32_t)grid > n_tiles) grid = (int)n_tiles;
  const size_t smem = sizeof(Smem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_send_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e - 1000;
  gemm_send_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma, mb, mc, g);
  return (int)cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) void rn_preload_gemm() {
  cudaFuncAttributes at;
  cudaFuncGetAttributes(&at, gemm_send_kernel);
  cudaFuncSetAttribute(gemm_send_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Smem) + 1024));
  cudaFuncGetAttributes(&at, gemm_send2_kernel);
  cudaFuncSetAttribute(gemm_send2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Smem2) + 1024));
  cudaFuncGetAttributes(&at, gemm_send3_kernel);
  cudaFuncSetAttribute(gemm_send3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Smem3) + 1024));
  encode_tiled();
}
