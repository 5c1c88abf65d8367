// N2 real-NIC backend: ConnectX through libibverbs + mlx5dv, the wire the softhca stands in for.
//
// Compiled two ways:
//   * rdma-core headers present (<infiniband/verbs.h>, <infiniband/mlx5dv.h>): the full path --
//     device open, PD/CQ/RC-QP, MR on host memory, on GPU HBM through nvidia-peermem (plain
//     ibv_reg_mr on the device pointer: the reference's whole purpose, README.md:5-6) or through a
//     dma-buf fd (ibv_reg_dmabuf_mr), loopback connect, host-posted WRITE/READ/SEND (the baselines
//     B0-B2 of BASELINE.md), and mlx5dv_init_obj() to expose the raw SQ / doorbell record / BlueFlame
//     register of a QP so the SAME device-side posting code (hca/post.cuh) can drive the NIC (K1).
//     Libraries are dlopen()ed, so the build never links against them.
//   * headers absent (this image: no rdma-core, and the GPU box exposes no /dev/infiniband -- gpurun
//     probe in DESIGN.md): only rn_verbs_available() / rn_verbs_why() exist and say so.  Nothing above
//     this file changes: Context picks the softhca wire.
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#define RN_API extern "C" __attribute__((visibility("default")))

#if defined(__has_include)
#if __has_include(<infiniband/verbs.h>) && __has_include(<infiniband/mlx5dv.h>)
#define RN_HAVE_VERBS_HEADERS 1
#endif
#endif

static char g_why[256] = "";

static bool uverbs_nodes_present() { return access("/dev/infiniband", R_OK | X_OK) == 0; }

#ifndef RN_HAVE_VERBS_HEADERS

RN_API int rn_verbs_available() {
  snprintf(g_why, sizeof g_why, "built without rdma-core headers; runtime: libibverbs %s, /dev/infiniband %s",
           dlopen("libibverbs.so.1", RTLD_LAZY | RTLD_LOCAL) ? "loadable" : "not installed",
           uverbs_nodes_present() ? "present" : "absent");
  return 0;
}
RN_API const char* rn_verbs_why() { return g_why; }
RN_API int rn_verbs_compiled() { return 0; }

#else  // ------------------------------------------------------------------ full backend

#include <cuda_runtime.h>
#include <errno.h>
#include <infiniband/mlx5dv.h>
#include <infiniband/verbs.h>

namespace {
struct Api {
  void* verbs = nullptr;
  void* mlx5 = nullptr;
  struct ibv_device** (*get_device_list)(int*) = nullptr;
  void (*free_device_list)(struct ibv_device**) = nullptr;
  const char* (*get_device_name)(struct ibv_device*) = nullptr;
  struct ibv_context* (*open_device)(struct ibv_device*) = nullptr;
  int (*close_device)(struct ibv_context*) = nullptr;
  struct ibv_pd* (*alloc_pd)(struct ibv_context*) = nullptr;
  int (*dealloc_pd)(struct ibv_pd*) = nullptr;
  struct ibv_mr* (*reg_mr)(struct ibv_pd*, void*, size_t, int) = nullptr;             // ibv_reg_mr is a macro over reg_mr_iova2
  struct ibv_mr* (*reg_dmabuf_mr)(struct ibv_pd*, uint64_t, size_t, uint64_t, int, int) = nullptr;
  int (*dereg_mr)(struct ibv_mr*) = nullptr;
  struct ibv_cq* (*create_cq)(struct ibv_context*, int, void*, struct ibv_comp_channel*, int) = nullptr;
  int (*destroy_cq)(struct ibv_cq*) = nullptr;
  struct ibv_qp* (*create_qp)(struct ibv_pd*, struct ibv_qp_init_attr*) = nullptr;
  int (*destroy_qp)(struct ibv_qp*) = nullptr;
  int (*modify_qp)(struct ibv_qp*, struct ibv_qp_attr*, int) = nullptr;
  int (*query_port)(struct ibv_context*, uint8_t, struct ibv_port_attr*) = nullptr;
  int (*query_gid)(struct ibv_context*, uint8_t, int, union ibv_gid*) = nullptr;
  int (*dv_init_obj)(struct mlx5dv_obj*, uint64_t) = nullptr;
  bool ok = false;
};
Api& api() {
  static Api a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  a.verbs = dlopen("libibverbs.so.1", RTLD_NOW | RTLD_GLOBAL);
  a.mlx5 = dlopen("libmlx5.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!a.verbs) { snprintf(g_why, sizeof g_why, "libibverbs.so.1: %s", dlerror()); return a; }
#define SYM(field, name) a.field = (decltype(a.field))dlsym(a.verbs, name)
  SYM(get_device_list, "ibv_get_device_list"); SYM(free_device_list, "ibv_free_device_list");
  SYM(get_device_name, "ibv_get_device_name"); SYM(open_device, "ibv_open_device"); SYM(close_device, "ibv_close_device");
  SYM(alloc_pd, "ibv_alloc_pd"); SYM(dealloc_pd, "ibv_dealloc_pd"); SYM(reg_mr, "ibv_reg_mr");
  SYM(reg_dmabuf_mr, "ibv_reg_dmabuf_mr"); SYM(dereg_mr, "ibv_dereg_mr"); SYM(create_cq, "ibv_create_cq");
  SYM(destroy_cq, "ibv_destroy_cq"); SYM(create_qp, "ibv_create_qp"); SYM(destroy_qp, "ibv_destroy_qp");
  SYM(modify_qp, "ibv_modify_qp"); SYM(query_port, "ibv_query_port"); SYM(query_gid, "ibv_query_gid");
#undef SYM
  if (a.mlx5) a.dv_init_obj = (decltype(a.dv_init_obj))dlsym(a.mlx5, "mlx5dv_init_obj");
  a.ok = a.get_device_list && a.open_device && a.alloc_pd && a.reg_mr && a.create_cq && a.create_qp && a.modify_qp;
  if (!a.ok) snprintf(g_why, sizeof g_why, "libibverbs is missing expected symbols");
  return a;
}

struct Dev {
  struct ibv_context* ctx = nullptr;
  struct ibv_pd* pd = nullptr;
  uint8_t port = 1;
  int gid_index = 0;
  struct ibv_port_attr pattr;
  union ibv_gid gid;
};
struct QpH {
  Dev* dev;
  struct ibv_cq* cq;
  struct ibv_qp* qp;
};
}  // namespace

RN_API int rn_verbs_compiled() { return 1; }
RN_API const char* rn_verbs_why() { return g_why; }
RN_API int rn_verbs_available() {
  Api& a = api();
  if (!a.ok) return 0;
  if (!uverbs_nodes_present()) { snprintf(g_why, sizeof g_why, "/dev/infiniband is not exposed to this container"); return 0; }
  int n = 0;
  struct ibv_device** l = a.get_device_list(&n);
  if (l) a.free_device_list(l);
  if (n <= 0) { snprintf(g_why, sizeof g_why, "no RDMA devices"); return 0; }
  return n;
}

RN_API int rn_verbs_device_name(int i, char* out, int cap) {
  Api& a = api();
  int n = 0;
  struct ibv_device** l = a.ok ? a.get_device_list(&n) : nullptr;
  if (!l || i >= n) { if (l) a.free_device_list(l); return -19; }
  snprintf(out, cap, "%s", a.get_device_name(l[i]));
  a.free_device_list(l);
  return 0;
}

RN_API void* rn_verbs_open(const char* name, int port, int gid_index) {
  Api& a = api();
  if (!a.ok) return nullptr;
  int n = 0;
  struct ibv_device** l = a.get_device_list(&n);
  Dev* d = nullptr;
  for (int i = 0; l && i < n; ++i) {
    if (name && *name && strcmp(a.get_device_name(l[i]), name)) continue;
    struct ibv_context* c = a.open_device(l[i]);
    if (!c) continue;
    d = new Dev();
    d->ctx = c; d->port = (uint8_t)port; d->gid_index = gid_index;
    d->pd = a.alloc_pd(c);
    if (!d->pd || a.query_port(c, d->port, &d->pattr)) { if (d->pd) a.dealloc_pd(d->pd); a.close_device(c); delete d; d = nullptr; continue; }
    memset(&d->gid, 0, sizeof d->gid);
    a.query_gid(c, d->port, gid_index, &d->gid);
    break;
  }
  if (l) a.free_device_list(l);
  if (!d) snprintf(g_why, sizeof g_why, "cannot open RDMA device %s", name ? name : "(any)");
  return d;
}
RN_API int rn_verbs_close(void* dev) {
  Dev* d = (Dev*)dev;
  if (!d) return 0;
  api().dealloc_pd(d->pd);
  api().close_device(d->ctx);
  delete d;
  return 0;
}
RN_API int rn_verbs_port_active(void* dev) { return ((Dev*)dev)->pattr.state == IBV_PORT_ACTIVE; }
RN_API int rn_verbs_link_layer(void* dev) { return ((Dev*)dev)->pattr.link_layer; }

// mode 0: ibv_reg_mr on the pointer (host memory, or GPU HBM through nvidia-peermem / b200p2p)
// mode 1: ibv_reg_dmabuf_mr on an exported dma-buf fd
RN_API void* rn_verbs_reg_mr(void* dev, uint64_t ptr, uint64_t len, int mode, int dmabuf_fd, uint32_t* lkey, uint32_t* rkey) {
  Dev* d = (Dev*)dev;
  const int acc = IBV_ACCESS_LOCAL_WRITE | IBV_ACCESS_REMOTE_WRITE | IBV_ACCESS_REMOTE_READ;
  struct ibv_mr* mr = nullptr;
  if (mode == 1 && api().reg_dmabuf_mr) mr = api().reg_dmabuf_mr(d->pd, 0, len, ptr, dmabuf_fd, acc);
  else mr = api().reg_mr(d->pd, (void*)ptr, len, acc);
  if (!mr) { snprintf(g_why, sizeof g_why, "reg_mr(mode %d) failed: errno %d", mode, errno); return nullptr; }
  *lkey = mr->lkey; *rkey = mr->rkey;
  return mr;
}
RN_API int rn_verbs_dereg_mr(void* mr) { return api().dereg_mr((struct ibv_mr*)mr); }

RN_API void* rn_verbs_create_qp(void* dev, int depth) {
  Dev* d = (Dev*)dev;
  QpH* q = new QpH();
  q->dev = d;
  q->cq = api().create_cq(d->ctx, depth * 2, nullptr, nullptr, 0);
  struct ibv_qp_init_attr ia;
  memset(&ia, 0, sizeof ia);
  ia.send_cq = ia.recv_cq = q->cq;
  ia.qp_type = IBV_QPT_RC;
  ia.cap.max_send_wr = depth; ia.cap.max_recv_wr = depth; ia.cap.max_send_sge = 1; ia.cap.max_recv_sge = 1;
  q->qp = q->cq ? api().create_qp(d->pd, &ia) : nullptr;
  if (!q->qp) { if (q->cq) api().destroy_cq(q->cq); delete q; snprintf(g_why, sizeof g_why, "create_qp failed: errno %d", errno); return nullptr; }
  return q;
}
RN_API uint32_t rn_verbs_qpn(void* qp) { return ((QpH*)qp)->qp->qp_num; }

// RESET -> INIT -> RTR -> RTS towards (remote_qpn, remote_lid / gid); loopback when it names ourselves.
RN_API int rn_verbs_connect(void* qp, uint32_t remote_qpn, uint16_t remote_lid, const uint8_t* remote_gid16) {
  QpH* q = (QpH*)qp;
  Dev* d = q->dev;
  struct ibv_qp_attr a;
  memset(&a, 0, sizeof a);
  a.qp_state = IBV_QPS_INIT; a.pkey_index = 0; a.port_num = d->port;
  a.qp_access_flags = IBV_ACCESS_REMOTE_WRITE | IBV_ACCESS_REMOTE_READ | IBV_ACCESS_LOCAL_WRITE;
  int rc = api().modify_qp(q->qp, &a, IBV_QP_STATE | IBV_QP_PKEY_INDEX | IBV_QP_PORT | IBV_QP_ACCESS_FLAGS);
  if (rc) return -rc;
  memset(&a, 0, sizeof a);
  a.qp_state = IBV_QPS_RTR; a.path_mtu = IBV_MTU_4096; a.dest_qp_num = remote_qpn; a.rq_psn = 0;
  a.max_dest_rd_atomic = 16; a.min_rnr_timer = 12;
  a.ah_attr.dlid = remote_lid; a.ah_attr.port_num = d->port;
  if (d->pattr.link_layer == IBV_LINK_LAYER_ETHERNET || remote_gid16) {
    a.ah_attr.is_global = 1; a.ah_attr.grh.hop_limit = 1; a.ah_attr.grh.sgid_index = d->gid_index;
    memcpy(&a.ah_attr.grh.dgid, remote_gid16 ? remote_gid16 : d->gid.raw, 16);
  }
  rc = api().modify_qp(q->qp, &a, IBV_QP_STATE | IBV_QP_AV | IBV_QP_PATH_MTU | IBV_QP_DEST_QPN | IBV_QP_RQ_PSN |
                                    IBV_QP_MAX_DEST_RD_ATOMIC | IBV_QP_MIN_RNR_TIMER);
  if (rc) return -rc;
  memset(&a, 0, sizeof a);
  a.qp_state = IBV_QPS_RTS; a.timeout = 14; a.retry_cnt = 7; a.rnr_retry = 7; a.sq_psn = 0; a.max_rd_atomic = 16;
  rc = api().modify_qp(q->qp, &a, IBV_QP_STATE | IBV_QP_TIMEOUT | IBV_QP_RETRY_CNT | IBV_QP_RNR_RETRY | IBV_QP_SQ_PSN | IBV_QP_MAX_QP_RD_ATOMIC);
  return rc ? -rc : 0;
}
RN_API int rn_verbs_local_addr(void* qp, uint16_t* lid, uint8_t* gid16) {
  Dev* d = ((QpH*)qp)->dev;
  *lid = d->pattr.lid;
  memcpy(gid16, d->gid.raw, 16);
  return 0;
}

// Host-posted verbs: the "ib_write_bw on a peermem MR" baseline.
RN_API int rn_verbs_post(void* qp, int opcode, uint64_t laddr, uint32_t lkey, uint64_t raddr, uint32_t rkey, uint32_t len, int signaled) {
  QpH* q = (QpH*)qp;
  struct ibv_sge sge = {laddr, len, lkey};
  struct ibv_send_wr wr, *bad = nullptr;
  memset(&wr, 0, sizeof wr);
  wr.sg_list = &sge; wr.num_sge = 1;
  wr.opcode = opcode == 0x10 ? IBV_WR_RDMA_READ : (opcode == 0x0a ? IBV_WR_SEND : IBV_WR_RDMA_WRITE);
  wr.send_flags = signaled ? IBV_SEND_SIGNALED : 0;
  wr.wr.rdma.remote_addr = raddr; wr.wr.rdma.rkey = rkey;
  return -ibv_post_send(q->qp, &wr, &bad);
}
RN_API int rn_verbs_poll(void* qp, int max, int* statuses) {
  QpH* q = (QpH*)qp;
  struct ibv_wc wc[16];
  if (max > 16) max = 16;
  int n = ibv_poll_cq(q->cq, max, wc);
  for (int i = 0; i < n; ++i) statuses[i] = wc[i].status;
  return n;
}

// Raw queue geometry for GPU-initiated posting: the caller cudaHostRegister()s these (the BlueFlame
// page with cudaHostRegisterIoMemory) and builds a QpDev around them.
struct RnRawQp {
  uint64_t sq_buf; uint32_t sq_wqe_cnt, sq_stride; uint64_t dbrec; uint64_t bf_reg; uint32_t bf_size;
  uint64_t cq_buf; uint32_t cq_cqe_cnt, cq_cqe_size; uint64_t cq_dbrec; uint32_t qpn, cqn;
};
RN_API int rn_verbs_raw_qp(void* qp, RnRawQp* out) {
  QpH* q = (QpH*)qp;
  if (!api().dv_init_obj) return -38;
  struct mlx5dv_qp dq; struct mlx5dv_cq dc; struct mlx5dv_obj obj;
  memset(&dq, 0, sizeof dq); memset(&dc, 0, sizeof dc); memset(&obj, 0, sizeof obj);
  obj.qp.in = q->qp; obj.qp.out = &dq; obj.cq.in = q->cq; obj.cq.out = &dc;
  int rc = api().dv_init_obj(&obj, MLX5DV_OBJ_QP | MLX5DV_OBJ_CQ);
  if (rc) return -rc;
  out->sq_buf = (uint64_t)dq.sq.buf; out->sq_wqe_cnt = dq.sq.wqe_cnt; out->sq_stride = dq.sq.stride;
  out->dbrec = (uint64_t)dq.dbrec; out->bf_reg = (uint64_t)dq.bf.reg; out->bf_size = dq.bf.size;
  out->cq_buf = (uint64_t)dc.buf; out->cq_cqe_cnt = dc.cqe_cnt; out->cq_cqe_size = dc.cqe_size; out->cq_dbrec = (uint64_t)dc.dbrec;
  out->qpn = q->qp->qp_num; out->cqn = dc.cqn;
  return 0;
}
// N3 on real hardware: make the NIC's queues visible to the GPU so hca/post.cuh can drive it.
//   * SQ buffer and doorbell record are host memory owned by rdma-core: cudaHostRegister(Mapped) gives the
//     GPU a device pointer to the same bytes (the NIC keeps reading them where it always did);
//   * the BlueFlame / doorbell register is a PCIe BAR page of the HCA: cudaHostRegisterIoMemory maps it so an
//     SM's 8-byte store becomes the MMIO doorbell write (needs the driver's PeerMappingOverride / IoMemory
//     support; SURVEY.md section 7.4 item 2 -- the fallback is a CPU proxy ringing it);
//   * the CQ buffer is registered the same way for the device-side poller.
// The caller (hca_host.cu) wraps these pointers in a QpDev/CqDev so write_rdma_wqe / sq_submit / cq_poll_once
// run unchanged: the wire format, the doorbell-record layout ([0] receive, [1] send) and the CQE owner-bit
// rule are the ones this project already uses.
struct RnGpuQp { uint64_t sq_dev, dbrec_dev, bf_dev, cq_dev, cq_dbrec_dev; uint32_t sq_wqe_cnt, cq_cqe_cnt, qpn, cqn; };
RN_API int rn_verbs_map_qp_to_gpu(void* qp, RnGpuQp* out) {
  RnRawQp raw;
  int rc = rn_verbs_raw_qp(qp, &raw);
  if (rc) return rc;
  if (raw.sq_stride != 64 || raw.cq_cqe_size != 64) { snprintf(g_why, sizeof g_why, "unexpected WQE/CQE stride %u/%u", raw.sq_stride, raw.cq_cqe_size); return -22; }
  auto map = [&](uint64_t host, size_t bytes, unsigned flags, uint64_t* dev) -> int {
    uint64_t page = host & ~4095ull;
    size_t len = ((host + bytes + 4095) & ~4095ull) - page;
    cudaError_t e = cudaHostRegister((void*)page, len, flags);
    if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered) { snprintf(g_why, sizeof g_why, "cudaHostRegister(0x%llx): %s", (unsigned long long)page, cudaGetErrorString(e)); cudaGetLastError(); return -5; }
    cudaGetLastError();
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, (void*)page, 0) != cudaSuccess) return -5;
    *dev = (uint64_t)d + (host - page);
    return 0;
  };
  if ((rc = map(raw.sq_buf, (size_t)raw.sq_wqe_cnt * 64, cudaHostRegisterMapped | cudaHostRegisterPortable, &out->sq_dev))) return rc;
  if ((rc = map(raw.dbrec, 8, cudaHostRegisterMapped | cudaHostRegisterPortable, &out->dbrec_dev))) return rc;
  if ((rc = map(raw.cq_buf, (size_t)raw.cq_cqe_cnt * 64, cudaHostRegisterMapped | cudaHostRegisterPortable, &out->cq_dev))) return rc;
  if ((rc = map(raw.cq_dbrec, 8, cudaHostRegisterMapped | cudaHostRegisterPortable, &out->cq_dbrec_dev))) return rc;
  if ((rc = map(raw.bf_reg, raw.bf_size ? raw.bf_size : 256, cudaHostRegisterIoMemory | cudaHostRegisterMapped | cudaHostRegisterPortable, &out->bf_dev))) return rc;
  out->sq_wqe_cnt = raw.sq_wqe_cnt; out->cq_cqe_cnt = raw.cq_cqe_cnt; out->qpn = raw.qpn; out->cqn = raw.cqn;
  return 0;
}

RN_API int rn_verbs_destroy_qp(void* qp) {
  QpH* q = (QpH*)qp;
  api().destroy_qp(q->qp);
  api().destroy_cq(q->cq);
  delete q;
  return 0;
}
#endif
