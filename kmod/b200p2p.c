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

#include <rdma/peer_mem.h>
#include "nv-p2p.h"

#define B200P2P_DRIVER_VERSION "1.0"
#define B200P2P_DRIVER_NAME "b200p2p"

MODULE_AUTHOR("rocnrdma_b200 authors");
MODULE_LICENSE("Dual MIT/GPL");
MODULE_DESCRIPTION("NVIDIA B200 P2P bridge driver for the PeerDirect interface");
MODULE_VERSION(B200P2P_DRIVER_VERSION);
MODULE_SOFTDEP("pre: nvidia ib_core");

#define MSG_DBG(fmt, args...) pr_debug(B200P2P_DRIVER_NAME ": " fmt, ##args)
#define MSG_INFO(fmt, args...) pr_info(B200P2P_DRIVER_NAME ": " fmt, ##args)
#define MSG_ERR(fmt, args...) pr_err(B200P2P_DRIVER_NAME ": " fmt, ##args)
#define MSG_WARN(fmt, args...) pr_warn(B200P2P_DRIVER_NAME ": " fmt, ##args)

/* NVIDIA pins GPU memory in 64 KiB pages; addresses and lengths must be aligned to that. */
#define GPU_PAGE_SHIFT 16
#define GPU_PAGE_SIZE (1ULL << GPU_PAGE_SHIFT)
#define GPU_PAGE_MASK (~(GPU_PAGE_SIZE - 1))

static invalidate_peer_memory ib_invalidate_callback;
static void *ib_reg_handle;

/* Observability the reference lacks (printk only: SURVEY.md section 5). */
static atomic64_t stat_acquired, stat_pinned, stat_mapped, stat_revoked, stat_released;

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

	struct mutex lock;
	enum b200_ctx_state state;
	int invalidating; /* the invalidate upcall is in progress on this context */
	int early_revoke; /* a free callback fired before the pin was recorded (state still ACQUIRED) */

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
	MSG_DBG("free_callback: ctx %p va 0x%llx size 0x%llx\n", ctx, (unsigned long long)ctx->va,
		(unsigned long long)ctx->size);

	mutex_lock(&ctx->lock);
	if (ctx->state != CTX_PINNED && ctx->state != CTX_MAPPED) {
		/* the pin this callback belongs to has not been recorded yet (probe pin in acquire, or
		 * get_pages still returning): remember it so get_pages does not publish a dead pin */
		if (ctx->state == CTX_ACQUIRED)
			ctx->early_revoke = 1;
		mutex_unlock(&ctx->lock);
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
	pt = ctx->page_table;
	ctx->dma_mapping = NULL;
	ctx->page_table = NULL;
	ctx->invalidating = 0;
	mutex_unlock(&ctx->lock);
	if (map)
		nvidia_p2p_free_dma_mapping(map);
	if (pt)
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
	pin_va = (u64)addr & GPU_PAGE_MASK;
	pin_size = (((u64)addr + size + GPU_PAGE_SIZE - 1) & GPU_PAGE_MASK) - pin_va;

	ctx = kzalloc(sizeof(*ctx), GFP_KERNEL);
	if (!ctx) {
		/* as in the reference (amdp2p.c:140-144): failure to allocate reads as "not ours" */
		MSG_ERR("acquire: cannot allocate a context\n");
		return 0;
	}
	mutex_init(&ctx->lock);
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
		mutex_destroy(&ctx->lock);
		kfree(ctx);
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
		mutex_destroy(&ctx->lock);
		kfree(ctx);
		return 0;
	}
	ret = nvidia_p2p_put_pages(0, 0, pin_va, probe);
	if (ret)
		MSG_WARN("acquire: dropping the probe pin failed (%d)\n", ret);

	*client_context = ctx;
	__module_get(THIS_MODULE); /* no rmmod while registrations are live */
	atomic64_inc(&stat_acquired);
	MSG_DBG("acquire: GPU address, ctx %p\n", ctx);
	return 1;
}

static int b200_get_pages(unsigned long addr, size_t size, int write, int force, struct sg_table *sg_head,
			  void *client_context, u64 core_context)
{
	struct b200_mem_context *ctx = client_context;
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

	ret = nvidia_p2p_get_pages(0, 0, ctx->pin_va, ctx->pin_size, &ctx->page_table, b200_free_callback, ctx);
	if (ret || !ctx->page_table) {
		MSG_ERR("get_pages: nvidia_p2p_get_pages failed: %d\n", ret);
		ctx->page_table = NULL;
		return ret ? ret : -EINVAL;
	}
	mutex_lock(&ctx->lock);
	if (ctx->early_revoke) {
		/* the memory was freed while the driver was still handing us the pin */
		struct nvidia_p2p_page_table *dead = ctx->page_table;

		ctx->page_table = NULL;
		ctx->early_revoke = 0;
		mutex_unlock(&ctx->lock);
		nvidia_p2p_free_page_table(dead);
		MSG_WARN("get_pages: range was freed during registration\n");
		return -EFAULT;
	}
	ctx->state = CTX_PINNED;
	mutex_unlock(&ctx->lock);
	atomic64_inc(&stat_pinned);
	/* as in the reference, sg_head is filled by dma_map (amdp2p.c:214) */
	return 0;
}

/* Map the pinned pages for ONE HCA and describe them in the scatterlist ib_core handed us. */
static int b200_dma_map(struct sg_table *sg_head, void *client_context, struct device *dma_device, int dmasync,
			int *nmap)
{
	struct b200_mem_context *ctx = client_context;
	struct nvidia_p2p_dma_mapping *map = NULL;
	struct scatterlist *sg;
	unsigned long psz;
	int ret, i;

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
	mutex_unlock(&ctx->lock);

	ret = nvidia_p2p_dma_map_pages(to_pci_dev(dma_device), ctx->page_table, &map);
	if (ret || !map) {
		MSG_ERR("dma_map: nvidia_p2p_dma_map_pages failed: %d\n", ret);
		return ret ? ret : -EINVAL;
	}
	ret = sg_alloc_table(sg_head, map->entries, GFP_KERNEL);
	if (ret) {
		nvidia_p2p_dma_unmap_pages(to_pci_dev(dma_device), ctx->page_table, map);
		return ret;
	}
	psz = page_size_of(ctx->page_table);
	for_each_sg(sg_head->sgl, sg, map->entries, i) {
		sg->offset = 0;
		sg->length = psz;
		sg_dma_address(sg) = map->dma_addresses[i];
		sg_dma_len(sg) = psz;
	}
	mutex_lock(&ctx->lock);
	if (ctx->state != CTX_PINNED) {
		/* revoked while we were mapping: the free callback owns the teardown of `map` only if it
		 * saw it, and it did not -- release it here */
		mutex_unlock(&ctx->lock);
		sg_free_table(sg_head);
		nvidia_p2p_free_dma_mapping(map);
		return -EINVAL;
	}
	ctx->dma_mapping = map;
	ctx->mapped_dev = to_pci_dev(dma_device);
	ctx->sg_allocated = 1;
	ctx->state = CTX_MAPPED;
	mutex_unlock(&ctx->lock);
	*nmap = map->entries;
	atomic64_inc(&stat_mapped);
	return 0;
}

static int b200_dma_unmap(struct sg_table *sg_head, void *client_context, struct device *dma_device)
{
	struct b200_mem_context *ctx = client_context;
	struct nvidia_p2p_dma_mapping *map = NULL;
	struct nvidia_p2p_page_table *pt = NULL;
	struct pci_dev *pdev = NULL;
	int revoked, free_sg;

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
	} else if (revoked && !ctx->invalidating) {
		pt = ctx->page_table;
		ctx->page_table = NULL;
	}
	mutex_unlock(&ctx->lock);

	if (!pt)
		return; /* never pinned, or the revoke path owns (or already did) the free */
	if (revoked) {
		nvidia_p2p_free_page_table(pt); /* NOT put_pages: the driver already unpinned */
	} else {
		ret = nvidia_p2p_put_pages(0, 0, ctx->pin_va, pt);
		if (ret)
			MSG_ERR("put_pages: nvidia_p2p_put_pages failed: %d\n", ret);
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

	if (!ctx)
		return;
	/* ib_core calls release last; be defensive about a teardown that skipped steps */
	if (ctx->state == CTX_MAPPED || ctx->state == CTX_PINNED || ctx->page_table || ctx->dma_mapping) {
		MSG_WARN("release: context %p still holds a pin, dropping it\n", ctx);
		b200_dma_unmap(NULL, ctx, NULL);
		b200_put_pages(NULL, ctx);
	}
	mutex_destroy(&ctx->lock);
	kfree(ctx);
	module_put(THIS_MODULE);
	atomic64_inc(&stat_released);
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

static int __init b200p2p_init(void)
{
	MSG_INFO("init (GPU page %llu KiB)\n", (unsigned long long)(GPU_PAGE_SIZE >> 10));
	strscpy(b200_mem_client.name, B200P2P_DRIVER_NAME, sizeof(b200_mem_client.name));
	strscpy(b200_mem_client.version, B200P2P_DRIVER_VERSION, sizeof(b200_mem_client.version));
	ib_reg_handle = ib_register_peer_memory_client(&b200_mem_client, &ib_invalidate_callback);
	if (!ib_reg_handle) {
		MSG_ERR("cannot register the peer memory client\n");
		return -EINVAL;
	}
	return 0;
}

/* Not reached while any registration is live: acquire holds a module reference. */
static void __exit b200p2p_exit(void)
{
	MSG_INFO("cleanup (acquired %ld pinned %ld mapped %ld revoked %ld released %ld)\n",
		 atomic64_read(&stat_acquired), atomic64_read(&stat_pinned), atomic64_read(&stat_mapped),
		 atomic64_read(&stat_revoked), atomic64_read(&stat_released));
	ib_unregister_peer_memory_client(ib_reg_handle);
	ib_reg_handle = NULL;
}

module_init(b200p2p_init);
module_exit(b200p2p_exit);
