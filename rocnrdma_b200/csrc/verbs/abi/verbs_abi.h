// The slice of the libibverbs userspace ABI this project uses, declared here so the ConnectX
// backend (verbs/verbs_dl.cc) and the in-tree mock provider (mockverbs/) compile in EVERY build --
// the image ships no rdma-core headers, and the backend only ever dlopen()s the libraries.
//
// Self-written from the stable ABI contract of libibverbs.so.1 (IBVERBS_1.x): only layouts and
// values that the library guarantees across releases are declared -- object headers whose leading
// fields applications read directly (ibv_mr::lkey, ibv_qp::qp_num, ...), the attribute structs passed
// by pointer, the enum values, and the fast-path dispatch through ibv_context::ops (post_send /
// post_recv / poll_cq are inline functions in the real header, not exported symbols).  Everything
// lives in namespace rnabi so that a build on a machine that does have <infiniband/verbs.h> cannot
// collide with it; the struct and field names follow the upstream ones so the code reads like verbs
// code.  tests/test_verbs_abi.py pins the sizes and offsets (x86-64 LP64).
//
// Reference parity: the reference only says "IB Verbs interface must be used" (README.md:67) and
// leaves the userspace to the application; this is the userspace.
#pragma once
#include <pthread.h>
#include <stddef.h>
#include <sys/types.h>
#include <stdint.h>

namespace rnabi {

typedef uint16_t be16_t;
typedef uint32_t be32_t;
typedef uint64_t be64_t;

// ------------------------------------------------------------------ enums
enum ibv_node_type { IBV_NODE_UNKNOWN = -1, IBV_NODE_CA = 1, IBV_NODE_SWITCH, IBV_NODE_ROUTER, IBV_NODE_RNIC };
enum ibv_transport_type { IBV_TRANSPORT_UNKNOWN = -1, IBV_TRANSPORT_IB = 0, IBV_TRANSPORT_IWARP };
enum ibv_port_state { IBV_PORT_NOP = 0, IBV_PORT_DOWN = 1, IBV_PORT_INIT = 2, IBV_PORT_ARMED = 3, IBV_PORT_ACTIVE = 4, IBV_PORT_ACTIVE_DEFER = 5 };
enum { IBV_LINK_LAYER_UNSPECIFIED = 0, IBV_LINK_LAYER_INFINIBAND = 1, IBV_LINK_LAYER_ETHERNET = 2 };
enum ibv_mtu { IBV_MTU_256 = 1, IBV_MTU_512 = 2, IBV_MTU_1024 = 3, IBV_MTU_2048 = 4, IBV_MTU_4096 = 5 };
enum ibv_access_flags {
  IBV_ACCESS_LOCAL_WRITE = 1, IBV_ACCESS_REMOTE_WRITE = 1 << 1, IBV_ACCESS_REMOTE_READ = 1 << 2,
  IBV_ACCESS_REMOTE_ATOMIC = 1 << 3, IBV_ACCESS_MW_BIND = 1 << 4, IBV_ACCESS_ZERO_BASED = 1 << 5,
  IBV_ACCESS_ON_DEMAND = 1 << 6, IBV_ACCESS_HUGETLB = 1 << 7, IBV_ACCESS_RELAXED_ORDERING = 1 << 20,
};
enum ibv_qp_type { IBV_QPT_RC = 2, IBV_QPT_UC = 3, IBV_QPT_UD = 4, IBV_QPT_RAW_PACKET = 8, IBV_QPT_XRC_SEND = 9, IBV_QPT_XRC_RECV = 10, IBV_QPT_DRIVER = 0xff };
enum ibv_qp_state { IBV_QPS_RESET = 0, IBV_QPS_INIT, IBV_QPS_RTR, IBV_QPS_RTS, IBV_QPS_SQD, IBV_QPS_SQE, IBV_QPS_ERR, IBV_QPS_UNKNOWN };
enum ibv_mig_state { IBV_MIG_MIGRATED = 0, IBV_MIG_REARM, IBV_MIG_ARMED };
enum ibv_qp_attr_mask {
  IBV_QP_STATE = 1 << 0, IBV_QP_CUR_STATE = 1 << 1, IBV_QP_EN_SQD_ASYNC_NOTIFY = 1 << 2, IBV_QP_ACCESS_FLAGS = 1 << 3,
  IBV_QP_PKEY_INDEX = 1 << 4, IBV_QP_PORT = 1 << 5, IBV_QP_QKEY = 1 << 6, IBV_QP_AV = 1 << 7, IBV_QP_PATH_MTU = 1 << 8,
  IBV_QP_TIMEOUT = 1 << 9, IBV_QP_RETRY_CNT = 1 << 10, IBV_QP_RNR_RETRY = 1 << 11, IBV_QP_RQ_PSN = 1 << 12,
  IBV_QP_MAX_QP_RD_ATOMIC = 1 << 13, IBV_QP_ALT_PATH = 1 << 14, IBV_QP_MIN_RNR_TIMER = 1 << 15, IBV_QP_SQ_PSN = 1 << 16,
  IBV_QP_MAX_DEST_RD_ATOMIC = 1 << 17, IBV_QP_PATH_MIG_STATE = 1 << 18, IBV_QP_CAP = 1 << 19, IBV_QP_DEST_QPN = 1 << 20,
  IBV_QP_RATE_LIMIT = 1 << 25,
};
enum ibv_wr_opcode {
  IBV_WR_RDMA_WRITE = 0, IBV_WR_RDMA_WRITE_WITH_IMM, IBV_WR_SEND, IBV_WR_SEND_WITH_IMM, IBV_WR_RDMA_READ,
  IBV_WR_ATOMIC_CMP_AND_SWP, IBV_WR_ATOMIC_FETCH_AND_ADD, IBV_WR_LOCAL_INV, IBV_WR_BIND_MW, IBV_WR_SEND_WITH_INV,
};
enum ibv_send_flags { IBV_SEND_FENCE = 1 << 0, IBV_SEND_SIGNALED = 1 << 1, IBV_SEND_SOLICITED = 1 << 2, IBV_SEND_INLINE = 1 << 3 };
enum ibv_wc_status {
  IBV_WC_SUCCESS = 0, IBV_WC_LOC_LEN_ERR, IBV_WC_LOC_QP_OP_ERR, IBV_WC_LOC_EEC_OP_ERR, IBV_WC_LOC_PROT_ERR,
  IBV_WC_WR_FLUSH_ERR, IBV_WC_MW_BIND_ERR, IBV_WC_BAD_RESP_ERR, IBV_WC_LOC_ACCESS_ERR, IBV_WC_REM_INV_REQ_ERR,
  IBV_WC_REM_ACCESS_ERR, IBV_WC_REM_OP_ERR, IBV_WC_RETRY_EXC_ERR, IBV_WC_RNR_RETRY_EXC_ERR, IBV_WC_LOC_RDD_VIOL_ERR,
  IBV_WC_REM_INV_RD_REQ_ERR, IBV_WC_REM_ABORT_ERR, IBV_WC_INV_EECN_ERR, IBV_WC_INV_EEC_STATE_ERR, IBV_WC_FATAL_ERR,
  IBV_WC_RESP_TIMEOUT_ERR, IBV_WC_GENERAL_ERR,
};
enum ibv_wc_opcode {
  IBV_WC_SEND = 0, IBV_WC_RDMA_WRITE, IBV_WC_RDMA_READ, IBV_WC_COMP_SWAP, IBV_WC_FETCH_ADD, IBV_WC_BIND_MW,
  IBV_WC_LOCAL_INV, IBV_WC_TSO, IBV_WC_RECV = 1 << 7, IBV_WC_RECV_RDMA_WITH_IMM,
};
enum ibv_wc_flags { IBV_WC_GRH = 1 << 0, IBV_WC_WITH_IMM = 1 << 1 };

// ------------------------------------------------------------------ objects (leading fields are ABI)
enum { IBV_SYSFS_NAME_MAX = 64, IBV_SYSFS_PATH_MAX = 256 };

struct ibv_context;
struct ibv_pd;
struct ibv_mr;
struct ibv_cq;
struct ibv_qp;
struct ibv_srq;
struct ibv_mw;
struct ibv_ah;
struct ibv_comp_channel;
struct ibv_send_wr;
struct ibv_recv_wr;
struct ibv_wc;
struct ibv_mw_bind;

struct _ibv_device_ops {
  struct ibv_context* (*_dummy1)(struct ibv_device*, int);
  void (*_dummy2)(struct ibv_context*);
};
struct ibv_device {
  struct _ibv_device_ops _ops;
  enum ibv_node_type node_type;
  enum ibv_transport_type transport_type;
  char name[IBV_SYSFS_NAME_MAX];        // kernel device name, e.g. "mlx5_0"
  char dev_name[IBV_SYSFS_NAME_MAX];    // uverbs device name, e.g. "uverbs0"
  char dev_path[IBV_SYSFS_PATH_MAX];
  char ibdev_path[IBV_SYSFS_PATH_MAX];
};

// Fast-path dispatch table embedded in every ibv_context.  The _compat_* slots are kept so the
// offsets of the live entries (poll_cq, post_send, post_recv, ...) never move.
struct ibv_context_ops {
  void* _compat_query_device;
  void* _compat_query_port;
  void* _compat_alloc_pd;
  void* _compat_dealloc_pd;
  void* _compat_reg_mr;
  void* _compat_rereg_mr;
  void* _compat_dereg_mr;
  struct ibv_mw* (*alloc_mw)(struct ibv_pd*, int);
  int (*bind_mw)(struct ibv_qp*, struct ibv_mw*, struct ibv_mw_bind*);
  int (*dealloc_mw)(struct ibv_mw*);
  void* _compat_create_cq;
  int (*poll_cq)(struct ibv_cq*, int, struct ibv_wc*);
  int (*req_notify_cq)(struct ibv_cq*, int);
  void* _compat_cq_event;
  void* _compat_resize_cq;
  void* _compat_destroy_cq;
  void* _compat_create_srq;
  void* _compat_modify_srq;
  void* _compat_query_srq;
  void* _compat_destroy_srq;
  int (*post_srq_recv)(struct ibv_srq*, struct ibv_recv_wr*, struct ibv_recv_wr**);
  void* _compat_create_qp;
  void* _compat_query_qp;
  void* _compat_modify_qp;
  void* _compat_destroy_qp;
  int (*post_send)(struct ibv_qp*, struct ibv_send_wr*, struct ibv_send_wr**);
  int (*post_recv)(struct ibv_qp*, struct ibv_recv_wr*, struct ibv_recv_wr**);
  void* _compat_create_ah;
  void* _compat_destroy_ah;
  void* _compat_attach_mcast;
  void* _compat_detach_mcast;
  void* _compat_async_event;
};

struct ibv_context {
  struct ibv_device* device;
  struct ibv_context_ops ops;
  int cmd_fd;
  int async_fd;
  int num_comp_vectors;
  pthread_mutex_t mutex;
  void* abi_compat;
};

struct ibv_pd {
  struct ibv_context* context;
  uint32_t handle;
};

struct ibv_mr {
  struct ibv_context* context;
  struct ibv_pd* pd;
  void* addr;
  size_t length;
  uint32_t handle;
  uint32_t lkey;
  uint32_t rkey;
};

struct ibv_cq {
  struct ibv_context* context;
  struct ibv_comp_channel* channel;
  void* cq_context;
  uint32_t handle;
  int cqe;
  pthread_mutex_t mutex;
  pthread_cond_t cond;
  uint32_t comp_events_completed;
  uint32_t async_events_completed;
};

struct ibv_qp {
  struct ibv_context* context;
  void* qp_context;
  struct ibv_pd* pd;
  struct ibv_cq* send_cq;
  struct ibv_cq* recv_cq;
  struct ibv_srq* srq;
  uint32_t handle;
  uint32_t qp_num;
  enum ibv_qp_state state;
  enum ibv_qp_type qp_type;
  pthread_mutex_t mutex;
  pthread_cond_t cond;
  uint32_t events_completed;
};

// ------------------------------------------------------------------ attributes
union ibv_gid {
  uint8_t raw[16];
  struct { be64_t subnet_prefix; be64_t interface_id; } global;
};

struct ibv_port_attr {
  enum ibv_port_state state;
  enum ibv_mtu max_mtu;
  enum ibv_mtu active_mtu;
  int gid_tbl_len;
  uint32_t port_cap_flags;
  uint32_t max_msg_sz;
  uint32_t bad_pkey_cntr;
  uint32_t qkey_viol_cntr;
  uint16_t pkey_tbl_len;
  uint16_t lid;
  uint16_t sm_lid;
  uint8_t lmc;
  uint8_t max_vl_num;
  uint8_t sm_sl;
  uint8_t subnet_timeout;
  uint8_t init_type_reply;
  uint8_t active_width;
  uint8_t active_speed;
  uint8_t phys_state;
  uint8_t link_layer;
  uint8_t flags;
  uint16_t port_cap_flags2;
  uint32_t active_speed_ex;
};

struct ibv_qp_cap {
  uint32_t max_send_wr;
  uint32_t max_recv_wr;
  uint32_t max_send_sge;
  uint32_t max_recv_sge;
  uint32_t max_inline_data;
};

struct ibv_qp_init_attr {
  void* qp_context;
  struct ibv_cq* send_cq;
  struct ibv_cq* recv_cq;
  struct ibv_srq* srq;
  struct ibv_qp_cap cap;
  enum ibv_qp_type qp_type;
  int sq_sig_all;
};

struct ibv_global_route {
  union ibv_gid dgid;
  uint32_t flow_label;
  uint8_t sgid_index;
  uint8_t hop_limit;
  uint8_t traffic_class;
};

struct ibv_ah_attr {
  struct ibv_global_route grh;
  uint16_t dlid;
  uint8_t sl;
  uint8_t src_path_bits;
  uint8_t static_rate;
  uint8_t is_global;
  uint8_t port_num;
};

struct ibv_qp_attr {
  enum ibv_qp_state qp_state;
  enum ibv_qp_state cur_qp_state;
  enum ibv_mtu path_mtu;
  enum ibv_mig_state path_mig_state;
  uint32_t qkey;
  uint32_t rq_psn;
  uint32_t sq_psn;
  uint32_t dest_qp_num;
  unsigned int qp_access_flags;
  struct ibv_qp_cap cap;
  struct ibv_ah_attr ah_attr;
  struct ibv_ah_attr alt_ah_attr;
  uint16_t pkey_index;
  uint16_t alt_pkey_index;
  uint8_t en_sqd_async_notify;
  uint8_t sq_draining;
  uint8_t max_rd_atomic;
  uint8_t max_dest_rd_atomic;
  uint8_t min_rnr_timer;
  uint8_t port_num;
  uint8_t timeout;
  uint8_t retry_cnt;
  uint8_t rnr_retry;
  uint8_t alt_port_num;
  uint8_t alt_timeout;
  uint32_t rate_limit;
};

// ------------------------------------------------------------------ work requests / completions
struct ibv_sge {
  uint64_t addr;
  uint32_t length;
  uint32_t lkey;
};

struct ibv_send_wr {
  uint64_t wr_id;
  struct ibv_send_wr* next;
  struct ibv_sge* sg_list;
  int num_sge;
  enum ibv_wr_opcode opcode;
  unsigned int send_flags;
  union { be32_t imm_data; uint32_t invalidate_rkey; };
  union {
    struct { uint64_t remote_addr; uint32_t rkey; } rdma;
    struct { uint64_t remote_addr; uint64_t compare_add; uint64_t swap; uint32_t rkey; } atomic;
    struct { struct ibv_ah* ah; uint32_t remote_qpn; uint32_t remote_qkey; } ud;
  } wr;
  union { struct { uint32_t remote_srqn; } xrc; } qp_type;
  union {
    struct { struct ibv_mw* mw; uint32_t rkey; struct { struct ibv_mr* mr; uint64_t addr; uint64_t length; unsigned int mw_access_flags; } bind_info; } bind_mw;
    struct { void* hdr; uint16_t hdr_sz; uint16_t mss; } tso;
  };
};

struct ibv_recv_wr {
  uint64_t wr_id;
  struct ibv_recv_wr* next;
  struct ibv_sge* sg_list;
  int num_sge;
};

struct ibv_wc {
  uint64_t wr_id;
  enum ibv_wc_status status;
  enum ibv_wc_opcode opcode;
  uint32_t vendor_err;
  uint32_t byte_len;
  union { be32_t imm_data; uint32_t invalidated_rkey; };
  uint32_t qp_num;
  uint32_t src_qp;
  unsigned int wc_flags;
  uint16_t pkey_index;
  uint16_t slid;
  uint8_t sl;
  uint8_t dlid_path_bits;
};

// ------------------------------------------------------------------ inline fast path (as in the real header)
static inline int ibv_poll_cq(struct ibv_cq* cq, int num_entries, struct ibv_wc* wc) {
  return cq->context->ops.poll_cq(cq, num_entries, wc);
}
static inline int ibv_post_send(struct ibv_qp* qp, struct ibv_send_wr* wr, struct ibv_send_wr** bad_wr) {
  return qp->context->ops.post_send(qp, wr, bad_wr);
}
static inline int ibv_post_recv(struct ibv_qp* qp, struct ibv_recv_wr* wr, struct ibv_recv_wr** bad_wr) {
  return qp->context->ops.post_recv(qp, wr, bad_wr);
}

// ------------------------------------------------------------------ mlx5 direct verbs (libmlx5.so.1, MLX5_1.x)
struct mlx5dv_qp {
  be32_t* dbrec;                                              // [0] receive, [1] send producer counters
  struct { void* buf; uint32_t wqe_cnt; uint32_t stride; } sq;
  struct { void* buf; uint32_t wqe_cnt; uint32_t stride; } rq;
  struct { void* reg; uint32_t size; } bf;                    // BlueFlame / doorbell register (a UAR page of the HCA)
  uint64_t comp_mask;
  off_t uar_mmap_offset;
  uint32_t tirn, tisn, rqn, sqn;
  uint64_t tir_icm_addr;
};
struct mlx5dv_cq {
  void* buf;
  be32_t* dbrec;                                              // [0] consumer index, [1] arm
  uint32_t cqe_cnt;
  uint32_t cqe_size;
  void* cq_uar;
  uint32_t cqn;
  uint64_t comp_mask;
};
struct mlx5dv_obj {
  struct { struct ibv_qp* in; struct mlx5dv_qp* out; } qp;
  struct { struct ibv_cq* in; struct mlx5dv_cq* out; } cq;
  struct { void* in; void* out; } srq;
  struct { void* in; void* out; } rwq;
  struct { void* in; void* out; } dm;
  struct { void* in; void* out; } ah;
  struct { void* in; void* out; } pd;
  struct { void* in; void* out; } devx;
};
enum mlx5dv_obj_type { MLX5DV_OBJ_QP = 1 << 0, MLX5DV_OBJ_CQ = 1 << 1, MLX5DV_OBJ_SRQ = 1 << 2, MLX5DV_OBJ_RWQ = 1 << 3, MLX5DV_OBJ_DM = 1 << 4, MLX5DV_OBJ_AH = 1 << 5, MLX5DV_OBJ_PD = 1 << 6 };

// The layouts above are load-bearing: a field that moves silently corrupts a real ibv_qp_attr.
static_assert(sizeof(void*) == 8, "LP64 only (the reference has the same restriction: README.md:65)");
static_assert(offsetof(ibv_device, name) == 24 && sizeof(ibv_device) == 24 + 64 + 64 + 256 + 256, "ibv_device");
static_assert(offsetof(ibv_context_ops, poll_cq) == 11 * 8 && offsetof(ibv_context_ops, post_send) == 25 * 8 &&
              offsetof(ibv_context_ops, post_recv) == 26 * 8 && sizeof(ibv_context_ops) == 32 * 8, "ibv_context_ops");
static_assert(offsetof(ibv_context, ops) == 8 && offsetof(ibv_context, cmd_fd) == 8 + 256, "ibv_context");
static_assert(offsetof(ibv_mr, lkey) == 36 && offsetof(ibv_mr, rkey) == 40, "ibv_mr");
static_assert(offsetof(ibv_qp, qp_num) == 52 && offsetof(ibv_qp, state) == 56, "ibv_qp");
static_assert(offsetof(ibv_cq, cqe) == 28, "ibv_cq");
static_assert(sizeof(ibv_sge) == 16 && sizeof(ibv_wc) == 48 && sizeof(ibv_recv_wr) == 32, "wr/wc");
static_assert(offsetof(ibv_send_wr, opcode) == 28 && offsetof(ibv_send_wr, wr) == 40 && sizeof(ibv_send_wr) == 128, "ibv_send_wr");
static_assert(sizeof(ibv_global_route) == 24 && sizeof(ibv_ah_attr) == 32, "ah_attr");
static_assert(offsetof(ibv_qp_attr, ah_attr) == 56 && offsetof(ibv_qp_attr, pkey_index) == 120 && sizeof(ibv_qp_attr) == 144, "ibv_qp_attr");
static_assert(sizeof(ibv_qp_init_attr) == 64, "ibv_qp_init_attr");
static_assert(offsetof(ibv_port_attr, lid) == 34 && offsetof(ibv_port_attr, link_layer) == 46, "ibv_port_attr");
static_assert(offsetof(mlx5dv_qp, sq) == 8 && offsetof(mlx5dv_qp, rq) == 24 && offsetof(mlx5dv_qp, bf) == 40 && offsetof(mlx5dv_qp, comp_mask) == 56, "mlx5dv_qp");
static_assert(offsetof(mlx5dv_cq, cqe_cnt) == 16 && offsetof(mlx5dv_cq, cqn) == 32, "mlx5dv_cq");

}  // namespace rnabi
