// SPDX-License-Identifier: GPL-2.0 OR MIT
/*
 * b200p2p - PeerDirect peer-memory client for NVIDIA Blackwell (B200) GPUs.
 *
 * Makes ibv_reg_mr() on a CUDA device pointer work: when ib_core cannot pin a virtual range it
 * asks every registered peer-memory client whether the range is theirs; this module answers for
 * GPU HBM by pinning it through the NVIDIA driver's P2P page-table interface (nv-p2p.h) and
 * handing the HCA the bus addresses of the GPU pages, after which the NIC DMAs to and from HBM
 * with no host bounce.  Capability-for-capability counterpart of AMD's amdp2p bridge
 * (/root/reference/amdp2p.c), rebuilt for the NVIDIA interface rather than ported:
 *
 *   reference (amdp2p.c)                         here
 *   ------------------------------------------   ---------------------------------------------
 *   amd_acquire :112-167, KFD is_gpu_address      b200_acquire: nv-p2p has no classifier, so a
 *                                                 probe pin of the 64 KiB-aligned range decides
 *   amd_get_pages :169-216                        b200_get_pages: nvidia_p2p_get_pages + revoke hook
 *   amd_dma_map :219-264 (shallow sg copy,        b200_dma_map: nvidia_p2p_dma_map_pages() PER HCA,
 *     assumes IOMMU off, ignores dma_device)        then an sg_table built from that mapping
 *   amd_dma_unmap :266-282 (no-op)                b200_dma_unmap: really unmaps / frees
 *   amd_put_pages :283-313 (ACCESS_ONCE flag)     b200_put_pages: state machine, see below
 *   amd_get_page_size :314-343 (4096 fallback)    b200_get_page_size: 64 KiB GPU pages
 *   amd_release :345-360                          b200_release
 *   free_callback :88-109                         b200_free_callback
 *   init/cleanup :374-408                         b200p2p_init / b200p2p_exit
 *
 * Deliberate differences (SURVEY.md sections 3.4 and 7.5): the revoke-vs-put race the reference
 * guards with a bare flag set AFTER the invalidate call is replaced by a per-registration state
 * machine under a mutex; the pid reference leak at :121 has no analogue (NVIDIA's API is keyed by
 * the calling context, not a struct pid); strcpy into the client name buffers becomes strscpy;
 * dma_map honours the HCA it is asked to map for; counters are exported for observability.
 */
#include <linux/module.h>
#include <linux/kernel.h>
#include <linux/slab.h>
#include <linux/mutex.h>
#include <linux/pci.h>
#include <linux/scatterlist.h>
#include <linux/errno.h>
#include <linux/kref.h>
#include <linux/debugfs.h>
#include <linux/seq_file.h>

#include <rdma/peer_mem.h>
#include "nv-p2p.h"

#define B200P2P_DRIVER_VERSION "1.0"
#define B200P2P_DRIVER_NAME "b200p2p"

MODULE_AUTHOR("rocnrdma_b200 authors");
MODULE_LICENSE("Dual MIT/GPL");
MODULE_DESCRIPTION("NVIDIA B200 P2P bridge driver for the PeerDirect interface");
MODULE_VERSION(B200P2P_DRIVER_VERSION);
MODULE_SOFTDEP("pre: nvidia ib_core");

/*
 * Run-time knobs (the reference has none: SURVEY.md section 5 "config").  /sys/module/b200p2p/parameters/
 *   debug       0: callback breadcrumbs only through dynamic debug (as the reference, README.md:60)
 *               1: also print them at INFO level -- for boxes whose kernel lacks CONFIG_DYNAMIC_DEBUG
 *   max_pin_mb  refuse to claim a range larger than this many MiB (0 = no limit): a guard rail for
 *               shared machines, acquire() then answers "not mine" and ibv_reg_mr() fails cleanly
 *   enable      0: stay registered with ib_core but claim nothing new (existing registrations keep working and are torn
 *               down normally) -- hands new registrations to another peer-memory client (stock nvidia-peermem) or
 *               switches GPU registration off without unloading; the nvidia-peermem "peerdirect_support"-style switch
 */
static int debug;
module_param(debug, int, 0644);
MODULE_PARM_DESC(debug, "1: print per-callback breadcrumbs at INFO level");
static unsigned long max_pin_mb;
module_param(max_pin_mb, ulong, 0644);
MODULE_PARM_DESC(max_pin_mb, "largest range (MiB) this client will claim, 0 = unlimited");
static int enable = 1;
module_param(enable, int, 0644);
MODULE_PARM_DESC(enable, "0: claim no new ranges (existing registrations are unaffected)");

#define MSG_DBG(fmt, args...)                                                       \
	do {                                                                        \
		if (debug)                                                          \
			pr_info(B200P2P_DRIVER_NAME ": " fmt, ##args);              \
		else                                                                \
			pr_debug(B200P2P_DRIVER_NAME ": " fmt, ##args);             \
	} while (0)
#define MSG_INFO(fmt, args...) pr_info(B200P2P_DRIVER_NAME ": " fmt, ##args)
#define MSG_ERR(fmt, args...) pr_err(B200P2P_DRIVER_NAME ": " fmt, ##args)
#define MSG_WARN(fmt, args...) pr_warn(B200P2P_DRIVER_NAME ": " fmt, ##args)

/* NVIDIA pins GPU memory in 64 KiB pages; addresses and lengths must be aligned to that. */
#define GPU_PAGE_SHIFT 16
#define GPU_PAGE_SIZE (1ULL << GPU_PAGE_SHIFT)
#define GPU_PAGE_MASK (~(GPU_PAGE_SIZE - 1))

static invalidate_peer_memory ib_invalidate_callback;
static void *ib_reg_handle;

/* Observability the reference lacks (printk only: SURVEY.md section 5): counters readable at any time in
 * debugfs (<debugfs>/b200p2p/stats), not just printed at rmmod. */
static atomic64_t stat_acquired, stat_pinned, stat_mapped, stat_revoked, stat_released, stat_refused;
static atomic64_t stat_live;
static struct dentry *b200_debugfs_dir;

/*
 * Lifetime of one registration (one ibv_reg_mr on a GPU range):
 *
 *   ACQUIRED --get_pages--> PINNED --dma_map--> MAPPED
 *       |                      |                  |
 *       |                      +---- free_callback (GPU memory going away) ----> REVOKED
 *       |                      |                  |                                |
 *       +------ release <------+--put_pages--<----+--dma_unmap            dma_unmap / put_pages only
 *                                                                          free bookkeeping: the
 *                                                                          driver already tore the
 *                                                                          pin down
 * Rules taken from the NVIDIA interface: after the free callback has fired, nvidia_p2p_put_pages()
 * and nvidia_p2p_dma_unmap_pages() must NOT be called for that pin; the page table and DMA mapping
 * are released with nvidia_p2p_free_page_table() / nvidia_p2p_free_dma_mapping() instead.
 */
enum b200_ctx_state {
	CTX_ACQUIRED = 0,
	CTX_PINNED,
	CTX_MAPPED,
	CTX_REVOKED,
};

struct b200_mem_context {
	u64 va;   /* range as registered by the application */
	u64 size;
	u64 pin_va; /* the same range widened to GPU page boundaries */
	u64 pin_size;

	/*
	 * Lifetime of the context itself.  ib_core owns one reference from acquire to release; the free
	 * callback takes its own for as long as it runs.  Without it a release that lands while the callback
	 * is inside (or just back from) the invalidate upcall -- ib_core tearing the MR down from the upcall,
	 * or a concurrent ibv_dereg_mr -- frees the context under the callback's feet.
	 */
	struct kref ref;
	struct mutex lock;
	enum b200_ctx_state state;
	int invalidating; /* the invalidate upcall is in progress on this context */
	int early_revoke; /* a free callback fired before the pin was recorded (state still ACQUIRED) */
	struct nvidia_p2p_page_table *zombie_pt; /* put_pages lost a race with a revoke (see b200_put_pages) */
	int pt_users;	  /* dma_map calls currently inside the NVIDIA driver with page_table (lock dropped) */
	int pt_orphaned;  /* the revoke path found pt_users != 0 and left the page table for the last user to free */

	struct nvidia_p2p_page_table *page_table;
	struct nvidia_p2p_dma_mapping *dma_mapping;
	struct pci_dev *mapped_dev;
	int sg_allocated;

	u64 core_context; /* ib_core's cookie for this MR */
};

static unsigned long page_size_of(const struct nvidia_p2p_page_table *pt)
{
	if (!pt)
		return GPU_PAGE_SIZE;
	switch (pt->page_size) {
	case NVIDIA_P2P_PAGE_SIZE_4KB:
		return 4096;
	case NVIDIA_P2P_PAGE_SIZE_128KB:
		return 128 * 1024;
	case NVIDIA_P2P_PAGE_SIZE_64KB:
	default:
		return GPU_PAGE_SIZE;
	}
}

static void b200_ctx_free(struct kref *ref)
{
	struct b200_mem_context *ctx = container_of(ref, struct b200_mem_context, ref);

	mutex_destroy(&ctx->lock);
	kfree(ctx);
	atomic64_dec(&stat_live);
	module_put(THIS_MODULE);
}

/*
 * Revocation: the NVIDIA driver is about to take the pinned pages away (cudaFree, process exit).
 * Ask ib_core to invalidate the MR -- it will re-enter dma_unmap/put_pages, possibly synchronously
 * on this very stack -- and make sure those paths no longer touch the pin.  The state flips to
 * REVOKED *before* the upcall, so a synchronous re-entry already sees it (the reference sets its
 * flag only after the upcall returns: amdp2p.c:103 then :108).
 */
static void b200_free_callback(void *data)
{
	struct b200_mem_context *ctx = data;
	struct nvidia_p2p_page_table *pt;
	struct nvidia_p2p_dma_mapping *map;
	u64 core_context;
	int upcall;

	if (!ctx) {
		MSG_WARN("free_callback: invalid client context\n");
		return;
	}
	/* The driver only calls back for a pin that exists, and a pin only exists between get_pages and
	 * put_pages -- i.e. while ib_core still holds its reference: taking ours here is safe, and from here
	 * on nothing ib_core does (release included) can free ctx before the final kref_put below. */
	kref_get(&ctx->ref);
	MSG_DBG("free_callback: ctx %p va 0x%llx size 0x%llx\n", ctx, (unsigned long long)ctx->va,
		(unsigned long long)ctx->size);

	mutex_lock(&ctx->lock);
	if (ctx->state != CTX_PINNED && ctx->state != CTX_MAPPED) {
		/* the pin this callback belongs to has not been recorded yet (probe pin in acquire, or
		 * get_pages still returning): remember it so get_pages does not publish a dead pin -- or it
		 * is the pin put_pages just failed to drop because this revoke was already under way: then
		 * its page table was parked for us to free */
		pt = NULL;
		if (ctx->state == CTX_ACQUIRED) {
			if (ctx->zombie_pt) {
				pt = ctx->zombie_pt;
				ctx->zombie_pt = NULL;
			} else {
				ctx->early_revoke = 1;
			}
		}
		mutex_unlock(&ctx->lock);
		if (pt) {
			nvidia_p2p_free_page_table(pt);
			kref_put(&ctx->ref, b200_ctx_free); /* the reference put_pages parked for this callback */
		}
		kref_put(&ctx->ref, b200_ctx_free);
		return;
	}
	ctx->state = CTX_REVOKED;
	ctx->invalidating = 1;
	core_context = ctx->core_context;
	upcall = ib_invalidate_callback != NULL;
	mutex_unlock(&ctx->lock);
	atomic64_inc(&stat_revoked);

	/* Not under the lock: ib_core re-enters our callbacks from here. */
	if (upcall)
		(*ib_invalidate_callback)(ib_reg_handle, core_context);

	/*
	 * Whatever ib_core did (synchronous teardown, deferred teardown, nothing yet), the driver's
	 * contract is that the bookkeeping objects are ours to free once this callback runs.
	 * dma_unmap/put_pages on a REVOKED context leave them alone while `invalidating` is set and
	 * only drop their own references, so there is exactly one owner for each free.
	 */
	mutex_lock(&ctx->lock);
	map = ctx->dma_mapping;
	ctx->dma_mapping = NULL;
	if (ctx->pt_users) {
		/* a dma_map is inside nvidia_p2p_dma_map_pages() with this page table right now (it had to drop
		 * the lock to call the driver): freeing it here would pull it from under that call.  The last
		 * such user frees it on its way out (b200_dma_map). */
		ctx->pt_orphaned = 1;
		pt = NULL;
	} else {
		pt = ctx->page_table;
		ctx->page_table = NULL;
	}
	ctx->invalidating = 0;
	mutex_unlock(&ctx->lock);
	if (map)
		nvidia_p2p_free_dma_mapping(map);
	if (pt)
		nvidia_p2p_free_page_table(pt);
	kref_put(&ctx->ref, b200_ctx_free); /* may be the last reference: release already ran */
}

/*
 * nvidia_p2p_put_pages() refused a pin we hold: a revoke of that very pin is in flight (its free callback has not
 * reached us yet, or ran a moment ago and found the context without a recorded pin).  The page table is still ours to
 * release -- with free_page_table, exactly once: here if the callback has already been, else by the callback when it
 * arrives.  In the second case the callback may come after ib_core has released the context, so a reference is
 * parked for it together with the page table.
 */
static void b200_put_refused(struct b200_mem_context *ctx, struct nvidia_p2p_page_table *pt, int ret, const char *who)
{
	int free_now;

	mutex_lock(&ctx->lock);
	free_now = ctx->early_revoke;
	ctx->early_revoke = 0;
	if (!free_now) {
		ctx->zombie_pt = pt;
		kref_get(&ctx->ref);
	}
	mutex_unlock(&ctx->lock);
	MSG_DBG("%s: lost a race with a revoke (%d), page table %s\n", who, ret,
		free_now ? "freed here" : "left to the free callback");
	if (free_now)
		nvidia_p2p_free_page_table(pt);
}

/* Ownership test.  Returns 1 (and a context) if [addr, addr+size) is GPU memory of the caller. */
static int b200_acquire(unsigned long addr, size_t size, void *peer_mem_private_data, char *peer_mem_name,
			void **client_context)
{
	struct b200_mem_context *ctx;
	struct nvidia_p2p_page_table *probe = NULL;
	u64 pin_va, pin_size;
	int ret;

	if (!size || !client_context)
		return 0;
	if (!enable) {
		MSG_DBG("acquire: disabled (enable=0), not claiming 0x%lx\n", addr);
		return 0;
	}
	pin_va = (u64)addr & GPU_PAGE_MASK;
	pin_size = (((u64)addr + size + GPU_PAGE_SIZE - 1) & GPU_PAGE_MASK) - pin_va;
	if (max_pin_mb && (pin_size >> 20) > max_pin_mb) {
		MSG_WARN("acquire: %llu MiB exceeds max_pin_mb=%lu, not claiming the range\n",
			 (unsigned long long)(pin_size >> 20), max_pin_mb);
		atomic64_inc(&stat_refused);
		return 0;
	}

	ctx = kzalloc(sizeof(*ctx), GFP_KERNEL);
	if (!ctx) {
		/* as in the reference (amdp2p.c:140-144): failure to allocate reads as "not ours" */
		MSG_ERR("acquire: cannot allocate a context\n");
		return 0;
	}
	mutex_init(&ctx->lock);
	kref_init(&ctx->ref);
	__module_get(THIS_MODULE); /* no rmmod while a context exists; dropped in b200_ctx_free */
	atomic64_inc(&stat_live);
	ctx->va = addr;
	ctx->size = size;
	ctx->pin_va = pin_va;
	ctx->pin_size = pin_size;
	ctx->state = CTX_ACQUIRED;

	/*
	 * nv-p2p offers no is_gpu_address(): the address is ours iff the driver agrees to pin it.
	 * The probe pin is dropped again immediately; get_pages() takes the real one.
	 */
	ret = nvidia_p2p_get_pages(0, 0, pin_va, pin_size, &probe, b200_free_callback, ctx);
	if (ret || !probe) {
		MSG_DBG("acquire: 0x%lx is not a GPU address (%d)\n", addr, ret);
		kref_put(&ctx->ref, b200_ctx_free);
		return 0;
	}
	/* A revoke that hits the probe pin finds state ACQUIRED: the callback only notes it. */
	mutex_lock(&ctx->lock);
	ret = ctx->early_revoke;
	ctx->early_revoke = 0;
	mutex_unlock(&ctx->lock);
	if (ret) {
		/* freed under our feet: the pin is already gone (put_pages would be a misuse), and the
		 * range is no longer GPU memory */
		nvidia_p2p_free_page_table(probe);
		kref_put(&ctx->ref, b200_ctx_free); /* the callback that noted the revoke may still hold its reference */
		return 0;
	}
	ret = nvidia_p2p_put_pages(0, 0, pin_va, probe);
	if (ret) {
		/* the memory is being freed right now: not (any longer) a GPU range worth claiming */
		b200_put_refused(ctx, probe, ret, "acquire");
		kref_put(&ctx->ref, b200_ctx_free);
		return 0;
	}

	*client_context = ctx;
	atomic64_inc(&stat_acquired);
	MSG_DBG("acquire: GPU address, ctx %p\n", ctx);
	return 1;
}

static int b200_get_pages(unsigned long addr, size_t size, int write, int force, struct sg_table *sg_head,
			  void *client_context, u64 core_context)
{
	struct b200_mem_context *ctx = client_context;
	struct nvidia_p2p_page_table *pt = NULL;
	int ret;

	if (!ctx) {
		MSG_WARN("get_pages: invalid client context\n");
		return -EINVAL;
	}
	if (addr != ctx->va) {
		MSG_WARN("get_pages: address 0x%lx differs from the acquired 0x%llx\n", addr,
			 (unsigned long long)ctx->va);
		return -EINVAL;
	}
	if (size != ctx->size) {
		MSG_WARN("get_pages: size 0x%zx differs from the acquired 0x%llx\n", size,
			 (unsigned long long)ctx->size);
		return -EINVAL;
	}
	mutex_lock(&ctx->lock);
	if (ctx->state != CTX_ACQUIRED) {
		mutex_unlock(&ctx->lock);
		MSG_ERR("get_pages: context is already pinned or revoked\n");
		return -EINVAL;
	}
	/* set before the pin exists: a revoke may arrive the instant the pin does */
	ctx->core_context = core_context;
	ctx->early_revoke = 0;
	mutex_unlock(&ctx->lock);

	ret = nvidia_p2p_get_pages(0, 0, ctx->pin_va, ctx->pin_size, &pt, b200_free_callback, ctx);
	if (ret || !pt) {
		MSG_ERR("get_pages: nvidia_p2p_get_pages failed: %d\n", ret);
		return ret ? ret : -EINVAL;
	}
	mutex_lock(&ctx->lock);
	if (ctx->early_revoke) {
		/* the memory was freed while the driver was still handing us the pin */
		ctx->early_revoke = 0;
		mutex_unlock(&ctx->lock);
		nvidia_p2p_free_page_table(pt);
		MSG_WARN("get_pages: range was freed during registration\n");
		return -EFAULT;
	}
	ctx->page_table = pt; /* published together with the state, under the lock the free callback takes */
	ctx->state = CTX_PINNED;
	mutex_unlock(&ctx->lock);
	atomic64_inc(&stat_pinned);
	/* as in the reference, sg_head is filled by dma_map (amdp2p.c:214) */
	return 0;
}

/* dma_map drops the context lock around the driver call, and a revoke that lands in that window may end with ib_core
 * releasing the context from inside the invalidate upcall: hold a reference for the duration of the call. */
static int b200_dma_map_locked_out(struct sg_table *sg_head, struct b200_mem_context *ctx, struct device *dma_device, int *nmap);
static int b200_dma_map(struct sg_table *sg_head, void *client_context, struct device *dma_device, int dmasync, int *nmap)
{
	struct b200_mem_context *ctx = client_context;
	int ret;

	if (!ctx || !sg_head || !nmap)
		return -EINVAL;
	kref_get(&ctx->ref);
	ret = b200_dma_map_locked_out(sg_head, ctx, dma_device, nmap);
	kref_put(&ctx->ref, b200_ctx_free);
	return ret;
}

/* Map the pinned pages for ONE HCA and describe them in the scatterlist ib_core handed us. */
static int b200_dma_map_locked_out(struct sg_table *sg_head, struct b200_mem_context *ctx, struct device *dma_device, int *nmap)
{
	struct nvidia_p2p_dma_mapping *map = NULL;
	struct nvidia_p2p_page_table *pt;
	struct scatterlist *sg;
	unsigned long psz;
	int ret, i, have_sg = 0;

	if (!ctx || !sg_head || !nmap)
		return -EINVAL;
	if (!dma_device || !dev_is_pci(dma_device)) {
		MSG_ERR("dma_map: the DMA device is not a PCI function\n");
		return -EINVAL;
	}
	mutex_lock(&ctx->lock);
	if (ctx->state != CTX_PINNED || !ctx->page_table) {
		mutex_unlock(&ctx->lock);
		MSG_ERR("dma_map: pages are not pinned\n");
		return -EINVAL;
	}
	/* The driver call below cannot run under ctx->lock (the free callback takes it, and the driver may
	 * hold its own locks across that callback): mark the page table as in use instead, so a revoke that
	 * lands meanwhile leaves it alive (pt_orphaned) until we are out of the driver. */
	pt = ctx->page_table;
	ctx->pt_users++;
	mutex_unlock(&ctx->lock);

	ret = nvidia_p2p_dma_map_pages(to_pci_dev(dma_device), pt, &map);
	if (!ret && map) {
		ret = sg_alloc_table(sg_head, map->entries, GFP_KERNEL);
		if (!ret) {
			have_sg = 1;
			psz = page_size_of(pt);
			for_each_sg(sg_head->sgl, sg, map->entries, i) {
				sg->offset = 0;
				sg->length = psz;
				sg_dma_address(sg) = map->dma_addresses[i];
				sg_dma_len(sg) = psz;
			}
		}
	} else if (!ret) {
		ret = -EINVAL;
	}
	mutex_lock(&ctx->lock);
	ctx->pt_users--;
	if (ctx->state != CTX_PINNED || ret) {
		/* failed, or revoked while we were mapping.  After a revoke the free callback owns the teardown of
		 * what it saw -- it never saw `map`, and if it found us inside the driver it left the page table
		 * to the last user as well. */
		const int revoked = ctx->state == CTX_REVOKED;
		struct nvidia_p2p_page_table *orphan = NULL;

		if (revoked && ctx->pt_orphaned && !ctx->pt_users) {
			orphan = ctx->page_table;
			ctx->page_table = NULL;
			ctx->pt_orphaned = 0;
		}
		mutex_unlock(&ctx->lock);
		if (have_sg)
			sg_free_table(sg_head);
		if (map) {
			if (revoked)
				nvidia_p2p_free_dma_mapping(map);
			else
				nvidia_p2p_dma_unmap_pages(to_pci_dev(dma_device), pt, map);
		}
		if (orphan)
			nvidia_p2p_free_page_table(orphan);
		if (ret)
			MSG_ERR("dma_map: mapping for the HCA failed: %d\n", ret);
		return ret ? ret : -EINVAL;
	}
	ctx->dma_mapping = map;
	ctx->mapped_dev = to_pci_dev(dma_device);
	ctx->sg_allocated = 1;
	ctx->state = CTX_MAPPED;
	*nmap = map->entries; /* before the unlock: once published, a revoke may free the mapping at any moment */
	mutex_unlock(&ctx->lock);
	atomic64_inc(&stat_mapped);
	return 0;
}

static int b200_dma_unmap(struct sg_table *sg_head, void *client_context, struct device *dma_device)
{
	struct b200_mem_context *ctx = client_context;
	struct nvidia_p2p_dma_mapping *map = NULL;
	struct nvidia_p2p_page_table *pt = NULL;
	struct pci_dev *pdev = NULL;
	int revoked, free_sg, pinned_pt = 0;

	if (!ctx)
		return -EINVAL;
	mutex_lock(&ctx->lock);
	revoked = ctx->state == CTX_REVOKED;
	free_sg = ctx->sg_allocated;
	ctx->sg_allocated = 0;
	if (ctx->state == CTX_MAPPED) {
		map = ctx->dma_mapping;
		pt = ctx->page_table;
		pdev = ctx->mapped_dev;
		ctx->dma_mapping = NULL;
		ctx->state = CTX_PINNED;
		ctx->pt_users++; /* the driver call below takes the page table with the lock dropped (as in dma_map) */
		pinned_pt = 1;
	} else if (revoked && !ctx->invalidating) {
		/* revoke already finished and nobody unmapped: the mapping object (if any) is ours */
		map = ctx->dma_mapping;
		ctx->dma_mapping = NULL;
	}
	mutex_unlock(&ctx->lock);

	if (free_sg && sg_head)
		sg_free_table(sg_head);
	if (map) {
		if (revoked)
			nvidia_p2p_free_dma_mapping(map);
		else
			nvidia_p2p_dma_unmap_pages(pdev, pt, map);
	}
	if (pinned_pt) {
		struct nvidia_p2p_page_table *orphan = NULL;

		mutex_lock(&ctx->lock);
		ctx->pt_users--;
		if (ctx->pt_orphaned && !ctx->pt_users) {
			/* a revoke came through while we were in the driver and left the page table to us */
			orphan = ctx->page_table;
			ctx->page_table = NULL;
			ctx->pt_orphaned = 0;
		}
		mutex_unlock(&ctx->lock);
		if (orphan)
			nvidia_p2p_free_page_table(orphan);
	}
	return 0;
}

static void b200_put_pages(struct sg_table *sg_head, void *client_context)
{
	struct b200_mem_context *ctx = client_context;
	struct nvidia_p2p_page_table *pt = NULL;
	int revoked, ret;

	if (!ctx)
		return;
	mutex_lock(&ctx->lock);
	revoked = ctx->state == CTX_REVOKED;
	if (ctx->state == CTX_MAPPED) {
		/* ib_core skipped dma_unmap: refuse to leak the mapping, but say so */
		mutex_unlock(&ctx->lock);
		MSG_WARN("put_pages: still DMA-mapped, unmapping first\n");
		b200_dma_unmap(sg_head, ctx, NULL);
		mutex_lock(&ctx->lock);
		revoked = ctx->state == CTX_REVOKED;
	}
	if (ctx->state == CTX_PINNED) {
		pt = ctx->page_table;
		ctx->page_table = NULL;
		ctx->state = CTX_ACQUIRED;
	} else if (revoked && !ctx->invalidating && !ctx->pt_orphaned) {
		pt = ctx->page_table;
		ctx->page_table = NULL;
	}
	mutex_unlock(&ctx->lock);

	if (!pt)
		return; /* never pinned, or the revoke path (or an in-flight dma_map) owns the free */
	if (revoked) {
		nvidia_p2p_free_page_table(pt); /* NOT put_pages: the driver already unpinned */
	} else {
		ret = nvidia_p2p_put_pages(0, 0, ctx->pin_va, pt);
		if (ret)
			b200_put_refused(ctx, pt, ret, "put_pages");
	}
}

static unsigned long b200_get_page_size(void *client_context)
{
	struct b200_mem_context *ctx = client_context;
	unsigned long psz;

	if (!ctx)
		return GPU_PAGE_SIZE;
	mutex_lock(&ctx->lock);
	psz = page_size_of(ctx->page_table); /* 64 KiB unless the driver says otherwise */
	mutex_unlock(&ctx->lock);
	return psz;
}

static void b200_release(void *client_context)
{
	struct b200_mem_context *ctx = client_context;
	int busy, leftover;

	if (!ctx)
		return;
	/* ib_core calls release last; be defensive about a teardown that skipped steps.  A revoke that is still
	 * in flight (invalidating) owns the pin's bookkeeping: leave it alone, it holds its own reference. */
	mutex_lock(&ctx->lock);
	busy = ctx->invalidating || ctx->pt_users;
	leftover = !busy && (ctx->state == CTX_MAPPED || ctx->state == CTX_PINNED || ctx->page_table || ctx->dma_mapping);
	mutex_unlock(&ctx->lock);
	if (leftover) {
		MSG_WARN("release: context %p still holds a pin, dropping it\n", ctx);
		b200_dma_unmap(NULL, ctx, NULL);
		b200_put_pages(NULL, ctx);
	}
	atomic64_inc(&stat_released);
	kref_put(&ctx->ref, b200_ctx_free); /* frees now, or when the in-flight free callback drops its reference */
}

static struct peer_memory_client b200_mem_client = {
	.acquire = b200_acquire,
	.get_pages = b200_get_pages,
	.dma_map = b200_dma_map,
	.dma_unmap = b200_dma_unmap,
	.put_pages = b200_put_pages,
	.get_page_size = b200_get_page_size,
	.release = b200_release,
};

static int b200_stats_show(struct seq_file *m, void *unused)
{
	seq_printf(m, "acquired %lld\npinned %lld\nmapped %lld\nrevoked %lld\nreleased %lld\nrefused %lld\nlive %lld\n",
		   (long long)atomic64_read(&stat_acquired), (long long)atomic64_read(&stat_pinned),
		   (long long)atomic64_read(&stat_mapped), (long long)atomic64_read(&stat_revoked),
		   (long long)atomic64_read(&stat_released), (long long)atomic64_read(&stat_refused),
		   (long long)atomic64_read(&stat_live));
	return 0;
}
DEFINE_SHOW_ATTRIBUTE(b200_stats);

static int __init b200p2p_init(void)
{
	MSG_INFO("init (GPU page %llu KiB, debug=%d, max_pin_mb=%lu, enable=%d)\n", (unsigned long long)(GPU_PAGE_SIZE >> 10), debug,
		 max_pin_mb, enable);
	atomic64_set(&stat_acquired, 0);
	atomic64_set(&stat_pinned, 0);
	atomic64_set(&stat_mapped, 0);
	atomic64_set(&stat_revoked, 0);
	atomic64_set(&stat_released, 0);
	atomic64_set(&stat_refused, 0);
	atomic64_set(&stat_live, 0);
	strscpy(b200_mem_client.name, B200P2P_DRIVER_NAME, sizeof(b200_mem_client.name));
	strscpy(b200_mem_client.version, B200P2P_DRIVER_VERSION, sizeof(b200_mem_client.version));
	ib_reg_handle = ib_register_peer_memory_client(&b200_mem_client, &ib_invalidate_callback);
	if (!ib_reg_handle) {
		MSG_ERR("cannot register the peer memory client\n");
		return -EINVAL;
	}
	/* debugfs is best effort: without it the module works, only the counters are not browsable */
	b200_debugfs_dir = debugfs_create_dir(B200P2P_DRIVER_NAME, NULL);
	if (!IS_ERR_OR_NULL(b200_debugfs_dir))
		debugfs_create_file("stats", 0444, b200_debugfs_dir, NULL, &b200_stats_fops);
	return 0;
}

/* Not reached while any registration is live: acquire holds a module reference. */
static void __exit b200p2p_exit(void)
{
	MSG_INFO("cleanup (acquired %lld pinned %lld mapped %lld revoked %lld released %lld)\n",
		 (long long)atomic64_read(&stat_acquired), (long long)atomic64_read(&stat_pinned),
		 (long long)atomic64_read(&stat_mapped), (long long)atomic64_read(&stat_revoked),
		 (long long)atomic64_read(&stat_released));
	debugfs_remove_recursive(b200_debugfs_dir);
	b200_debugfs_dir = NULL;
	ib_unregister_peer_memory_client(ib_reg_handle);
	ib_reg_handle = NULL;
}

module_init(b200p2p_init);
module_exit(b200p2p_exit);
