/* SPDX-License-Identifier: MIT */
/*
 * Simulation runtime: the two external layers of the kernel modules, in user space.
 *
 *   mock NVIDIA P2P provider   "GPU allocations" registered by the test, nvidia_p2p_get_pages with the
 *                              real interface's rules (64 KiB alignment, range must be GPU memory),
 *                              per-pin free callbacks fired by sim_gpu_free() (= cudaFree / process
 *                              exit), misuse detection (put_pages after revoke, double free) and
 *                              leak accounting for page tables and DMA mappings.
 *   mock ib_core PeerDirect    ib_register_peer_memory_client + a reg_mr/dereg_mr driver that calls the
 *                              client's callbacks in ib_core's order, and an invalidate callback that
 *                              tears the MR down either synchronously (re-entering the client from
 *                              inside its own free callback) or lazily at dereg -- both orders the
 *                              reference's bare-flag scheme could not tell apart.
 *   file / misc-device glue    open / ioctl / mmap / close entry points for the harness module.
 */
#include "sim_kernel.h"
#include "nv-p2p.h"
#include "rdma/peer_mem.h"

#define SIM_API __attribute__((visibility("default")))
#define GPU_PAGE 65536ULL
#define BUS_XOR 0x0000200000000000ULL /* bus address = va ^ BUS_XOR: recognisable, never equal to the va */

struct module sim_this_module = { {0}, "sim" };

/* ------------------------------------------------------------------ logging + allocation accounting */
static int g_log_count[4];
static int g_verbose;
static long g_live_allocs;
static int g_copy_fault_in = -1;
static int g_lock_errors;

void sim_log(int level, const char *fmt, ...)
{
	va_list ap;

	if (level >= 0 && level < 4)
		__atomic_add_fetch(&g_log_count[level], 1, __ATOMIC_RELAXED);
	if (!g_verbose)
		return;
	va_start(ap, fmt);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
}
void *sim_kzalloc(size_t n) { void *p = calloc(1, n ? n : 1); if (p) __sync_add_and_fetch(&g_live_allocs, 1); return p; }
void sim_kfree(const void *p) { if (p) { __sync_sub_and_fetch(&g_live_allocs, 1); free((void *)p); } }
int sim_mutex_lock(struct mutex *l)
{
	int rc = pthread_mutex_lock(&l->m);

	if (rc) { /* EDEADLK: the module tried to take a lock it already holds */
		__atomic_add_fetch(&g_lock_errors, 1, __ATOMIC_RELAXED);
		sim_log(0, "sim: mutex_lock error %d (recursive lock?)\n", rc);
	}
	return rc;
}
/* ------------------------------------------------------------------ debugfs / seq_file */
#define MAX_DBGFS 8
static struct { int live; char name[64]; const struct sim_seq_ops *fops; void *data; struct dentry d; } g_dbgfs[MAX_DBGFS];
static struct dentry g_dbgfs_root;
void seq_printf(struct seq_file *m, const char *fmt, ...)
{
	va_list ap;
	int n;

	va_start(ap, fmt);
	n = vsnprintf(m->buf + m->len, m->cap > m->len ? m->cap - m->len : 0, fmt, ap);
	va_end(ap);
	if (n > 0) m->len += (size_t)n < m->cap - m->len ? (size_t)n : m->cap - m->len;
}
struct dentry *debugfs_create_dir(const char *name, struct dentry *parent) { return &g_dbgfs_root; }
struct dentry *debugfs_create_file(const char *name, unsigned short mode, struct dentry *parent, void *data, const struct sim_seq_ops *fops)
{
	int i;

	for (i = 0; i < MAX_DBGFS; i++)
		if (!g_dbgfs[i].live) {
			g_dbgfs[i].live = 1; g_dbgfs[i].fops = fops; g_dbgfs[i].data = data;
			snprintf(g_dbgfs[i].name, sizeof g_dbgfs[i].name, "%s", name);
			return &g_dbgfs[i].d;
		}
	return NULL;
}
void debugfs_remove_recursive(struct dentry *d) { int i; if (d) for (i = 0; i < MAX_DBGFS; i++) g_dbgfs[i].live = 0; }
/* cat <debugfs>/.../<name>: returns the length, or -ENOENT */
SIM_API int sim_debugfs_read(const char *name, char *out, int cap)
{
	struct seq_file m = { out, (size_t)cap, 0, NULL };
	int i;

	for (i = 0; i < MAX_DBGFS; i++)
		if (g_dbgfs[i].live && !strcmp(g_dbgfs[i].name, name)) {
			m.private = g_dbgfs[i].data;
			if (cap > 0) out[0] = 0;
			g_dbgfs[i].fops->show(&m, NULL);
			return (int)m.len;
		}
	return -ENOENT;
}

size_t strscpy(char *dst, const char *src, size_t n)
{
	size_t l = strlen(src);

	if (!n) return 0;
	if (l >= n) l = n - 1;
	memcpy(dst, src, l);
	dst[l] = 0;
	return l;
}
unsigned long sim_copy(void *dst, const void *src, unsigned long n)
{
	if (g_copy_fault_in == 0) { g_copy_fault_in = -1; return n; }
	if (g_copy_fault_in > 0) g_copy_fault_in--;
	memcpy(dst, src, n);
	return 0;
}
int sg_alloc_table(struct sg_table *t, unsigned int nents, gfp_t gfp)
{
	t->sgl = sim_kzalloc(sizeof(struct scatterlist) * (nents ? nents : 1));
	if (!t->sgl) return -ENOMEM;
	t->nents = t->orig_nents = nents;
	return 0;
}
void sg_free_table(struct sg_table *t) { if (t && t->sgl) { sim_kfree(t->sgl); t->sgl = NULL; t->nents = t->orig_nents = 0; } }

SIM_API void sim_set_verbose(int v) { g_verbose = v; }
SIM_API int sim_log_count(int level) { return level >= 0 && level < 4 ? g_log_count[level] : -1; }
SIM_API void sim_log_reset(void) { memset(g_log_count, 0, sizeof g_log_count); }
SIM_API long sim_live_allocs(void) { return g_live_allocs; }
SIM_API void sim_set_copy_fault(int nth) { g_copy_fault_in = nth; }
SIM_API int sim_lock_errors(void) { return g_lock_errors; }
SIM_API int sim_module_refcount(void) { return atomic_read(&sim_this_module.refcnt); }

/* ------------------------------------------------------------------ mock NVIDIA P2P provider */
#define MAX_ALLOCS 64
#define MAX_PINS 256
struct gpu_alloc { u64 va, size; int live; };
struct pin {
	int live, revoked, cb_done;
	u64 va, size;
	struct nvidia_p2p_page_table *pt;
	void (*cb)(void *);
	void *cb_data;
};
static struct gpu_alloc g_allocs[MAX_ALLOCS];
static struct pin g_pins[MAX_PINS];
static pthread_mutex_t g_nv_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_live_pt, g_live_map, g_misuse, g_fail_next_get_pages, g_fail_next_dma_map;
static int g_revoke_during_get_pages; /* fire the free callback from inside get_pages (early revoke) */
static void (*g_dma_map_hook)(void);
static int g_concurrent;   /* multi-threaded stress: unpin / unmap calls may legitimately cross a revoke they could not have seen */

SIM_API int sim_gpu_alloc(u64 va, u64 size)
{
	int i;

	int rc = -ENOMEM;

	if ((va | size) & (GPU_PAGE - 1) || !size) return -EINVAL;
	pthread_mutex_lock(&g_nv_lock);
	for (i = 0; i < MAX_ALLOCS; i++)
		if (g_allocs[i].live && g_allocs[i].va == va && g_allocs[i].size == size) { rc = 0; break; }   /* already known */
	for (i = 0; rc && i < MAX_ALLOCS; i++)
		if (!g_allocs[i].live) { g_allocs[i].va = va; g_allocs[i].size = size; g_allocs[i].live = 1; rc = 0; }
	pthread_mutex_unlock(&g_nv_lock);
	return rc;
}
static int range_is_gpu(u64 va, u64 size)
{
	int i;

	for (i = 0; i < MAX_ALLOCS; i++)
		if (g_allocs[i].live && va >= g_allocs[i].va && va + size <= g_allocs[i].va + g_allocs[i].size) return 1;
	return 0;
}
SIM_API u64 sim_gpu_bus_addr(u64 va) { return va ^ BUS_XOR; }
/* With several threads a put_pages / dma_unmap_pages can be decided before, and arrive after, a revoke: the driver
 * refuses it and that is not a client bug.  The ordering rules are enforced by the single-threaded scenarios; the
 * stress keeps the checks that hold under any interleaving (unknown or double-freed page table, a freed page table
 * handed to the driver, leaks). */
SIM_API void sim_nv_set_concurrent(int on) { g_concurrent = on; }
/* Revoke allocation `va` from INSIDE the client's next nvidia_p2p_dma_map_pages() call (one shot): the window in which
 * the client has dropped its lock and handed the page table to the driver. */
static u64 g_hook_free_va;
SIM_API int sim_gpu_free(u64 va);
static void hook_free_once(void) { u64 va = g_hook_free_va; g_dma_map_hook = NULL; g_hook_free_va = 0; if (va) sim_gpu_free(va); }
SIM_API void sim_nv_revoke_during_dma_map(u64 va) { g_hook_free_va = va; g_dma_map_hook = va ? hook_free_once : NULL; }
SIM_API int sim_live_page_tables(void) { return g_live_pt; }
SIM_API int sim_live_dma_mappings(void) { return g_live_map; }
SIM_API int sim_nv_misuse(void) { return g_misuse; }
SIM_API void sim_nv_fail_next_get_pages(int n) { g_fail_next_get_pages = n; }
SIM_API void sim_nv_fail_next_dma_map(int n) { g_fail_next_dma_map = n; }
SIM_API void sim_nv_revoke_during_get_pages(int on) { g_revoke_during_get_pages = on; }
SIM_API int sim_live_pins(void) { int i, n = 0; for (i = 0; i < MAX_PINS; i++) n += g_pins[i].live && !g_pins[i].revoked; return n; }

int nvidia_p2p_get_pages(uint64_t token, uint32_t va_space, uint64_t va, uint64_t len,
			 struct nvidia_p2p_page_table **page_table, void (*free_callback)(void *), void *data)
{
	struct nvidia_p2p_page_table *pt;
	u32 n, i;
	int slot = -1;

	if (!page_table || !free_callback || !len || ((va | len) & (GPU_PAGE - 1))) return -EINVAL; /* as the real driver */
	if (g_fail_next_get_pages > 0 && --g_fail_next_get_pages == 0) return -ENOMEM;
	pthread_mutex_lock(&g_nv_lock);
	if (!range_is_gpu(va, len)) { pthread_mutex_unlock(&g_nv_lock); return -EINVAL; }
	for (i = 0; i < MAX_PINS; i++) if (!g_pins[i].live) { slot = (int)i; break; }
	if (slot < 0) { pthread_mutex_unlock(&g_nv_lock); return -ENOMEM; }
	n = (u32)(len / GPU_PAGE);
	pt = calloc(1, sizeof(*pt));
	pt->version = 0x00010002; pt->page_size = NVIDIA_P2P_PAGE_SIZE_64KB; pt->entries = n;
	pt->pages = calloc(n, sizeof(*pt->pages));
	for (i = 0; i < n; i++) {
		pt->pages[i] = calloc(1, sizeof(struct nvidia_p2p_page));
		pt->pages[i]->physical_address = sim_gpu_bus_addr(va + (u64)i * GPU_PAGE);
	}
	g_pins[slot] = (struct pin){ 1, 0, 0, va, len, pt, free_callback, data };
	g_live_pt++;
	*page_table = pt;
	pthread_mutex_unlock(&g_nv_lock);
	if (g_revoke_during_get_pages && free_callback) {
		/* the allocation dies while the caller is still inside get_pages */
		g_revoke_during_get_pages = 0;
		pthread_mutex_lock(&g_nv_lock);
		g_pins[slot].revoked = 1;
		pthread_mutex_unlock(&g_nv_lock);
		free_callback(data);
		pthread_mutex_lock(&g_nv_lock);
		if (g_pins[slot].live && g_pins[slot].pt == pt) g_pins[slot].cb_done = 1;
		pthread_mutex_unlock(&g_nv_lock);
	}
	return 0;
}
static struct pin *find_pin(struct nvidia_p2p_page_table *pt)
{
	int i;

	for (i = 0; i < MAX_PINS; i++) if (g_pins[i].live && g_pins[i].pt == pt) return &g_pins[i];
	return NULL;
}
static void destroy_pt(struct nvidia_p2p_page_table *pt)
{
	u32 i;

	for (i = 0; i < pt->entries; i++) free(pt->pages[i]);
	free(pt->pages);
	free(pt);
	g_live_pt--;
}
int nvidia_p2p_put_pages(uint64_t token, uint32_t va_space, uint64_t va, struct nvidia_p2p_page_table *pt)
{
	struct pin *p;

	pthread_mutex_lock(&g_nv_lock);
	p = find_pin(pt);
	if (!p) { g_misuse++; pthread_mutex_unlock(&g_nv_lock); sim_log(0, "nv-p2p: put_pages on an unknown page table\n"); return -EINVAL; }
	if (p->revoked) {
		/* Once the free callback has RETURNED the client knows the pin is gone: unpinning it is a bug.  While the
		 * revoke is still in flight the client could not know -- the driver just refuses, and the page table
		 * stays the client's to release with nvidia_p2p_free_page_table(). */
		if (p->cb_done && !g_concurrent) { g_misuse++; sim_log(0, "nv-p2p: put_pages after the free callback\n"); }
		pthread_mutex_unlock(&g_nv_lock);
		return -EINVAL;
	}
	if (p->va != va) { g_misuse++; pthread_mutex_unlock(&g_nv_lock); return -EINVAL; }
	p->live = 0;
	destroy_pt(pt);
	pthread_mutex_unlock(&g_nv_lock);
	return 0;
}
int nvidia_p2p_free_page_table(struct nvidia_p2p_page_table *pt)
{
	struct pin *p;

	pthread_mutex_lock(&g_nv_lock);
	p = find_pin(pt);
	if (!p) { g_misuse++; pthread_mutex_unlock(&g_nv_lock); sim_log(0, "nv-p2p: free_page_table on an unknown page table (double free?)\n"); return -EINVAL; }
	if (!p->revoked) { g_misuse++; sim_log(0, "nv-p2p: free_page_table on a live pin\n"); }
	p->live = 0;
	destroy_pt(pt);
	pthread_mutex_unlock(&g_nv_lock);
	return 0;
}
struct map_priv { int unmapped; };
int nvidia_p2p_dma_map_pages(struct pci_dev *peer, struct nvidia_p2p_page_table *pt, struct nvidia_p2p_dma_mapping **out)
{
	struct nvidia_p2p_dma_mapping *m;
	u32 i;

	if (!peer || !pt || !out) return -EINVAL;
	if (g_fail_next_dma_map > 0 && --g_fail_next_dma_map == 0) return -EIO;
	if (g_dma_map_hook) g_dma_map_hook();                 /* tests: something happens while the client is inside the driver */
	pthread_mutex_lock(&g_nv_lock);
	if (!find_pin(pt)) {                                    /* a freed page table reached the driver: the bug pt_users prevents */
		g_misuse++;
		pthread_mutex_unlock(&g_nv_lock);
		sim_log(0, "nv-p2p: dma_map_pages on a page table that no longer exists\n");
		return -EINVAL;
	}
	m = calloc(1, sizeof(*m));
	m->version = 0x00020003; m->page_size_type = NVIDIA_P2P_PAGE_SIZE_64KB; m->entries = pt->entries;
	m->dma_addresses = calloc(pt->entries, sizeof(u64));
	/* per-HCA IOVA: the device id is folded in so a test can tell two HCAs' mappings apart */
	for (i = 0; i < pt->entries; i++) m->dma_addresses[i] = pt->pages[i]->physical_address + ((u64)peer->dev.id << 52);
	m->pci_dev = peer;
	g_live_map++;
	*out = m;
	pthread_mutex_unlock(&g_nv_lock);
	return 0;
}
static void destroy_map(struct nvidia_p2p_dma_mapping *m) { free(m->dma_addresses); free(m); g_live_map--; }
int nvidia_p2p_dma_unmap_pages(struct pci_dev *peer, struct nvidia_p2p_page_table *pt, struct nvidia_p2p_dma_mapping *m)
{
	struct pin *p;

	if (!m) return -EINVAL;
	pthread_mutex_lock(&g_nv_lock);
	p = pt ? find_pin(pt) : NULL;
	if (p && p->revoked && p->cb_done && !g_concurrent) { g_misuse++; sim_log(0, "nv-p2p: dma_unmap_pages after the free callback\n"); }
	if (pt && !p) { g_misuse++; sim_log(0, "nv-p2p: dma_unmap_pages with a page table that no longer exists\n"); }
	if (peer != m->pci_dev) { g_misuse++; sim_log(0, "nv-p2p: dma_unmap_pages for a different device\n"); }
	destroy_map(m);
	pthread_mutex_unlock(&g_nv_lock);
	return 0;
}
int nvidia_p2p_free_dma_mapping(struct nvidia_p2p_dma_mapping *m)
{
	if (!m) return -EINVAL;
	pthread_mutex_lock(&g_nv_lock);
	destroy_map(m);
	pthread_mutex_unlock(&g_nv_lock);
	return 0;
}

/* cudaFree / process exit: every live pin overlapping the allocation is revoked (callbacks run unlocked). */
SIM_API int sim_gpu_free(u64 va)
{
	struct { void (*cb)(void *); void *data; int slot; struct nvidia_p2p_page_table *pt; } fire[MAX_PINS];
	int i, a = -1, n = 0;

	pthread_mutex_lock(&g_nv_lock);
	for (i = 0; i < MAX_ALLOCS; i++) if (g_allocs[i].live && g_allocs[i].va == va) { a = i; break; }
	if (a < 0) { pthread_mutex_unlock(&g_nv_lock); return -EINVAL; }
	for (i = 0; i < MAX_PINS; i++) {
		struct pin *p = &g_pins[i];

		if (!p->live || p->revoked) continue;
		if (p->va + p->size <= g_allocs[a].va || p->va >= g_allocs[a].va + g_allocs[a].size) continue;
		p->revoked = 1;
		if (p->cb) { fire[n].cb = p->cb; fire[n].data = p->cb_data; fire[n].slot = i; fire[n].pt = p->pt; n++; }
		else { /* a pin without a callback simply disappears with the memory */ p->live = 0; destroy_pt(p->pt); }
	}
	g_allocs[a].live = 0;
	pthread_mutex_unlock(&g_nv_lock);
	for (i = 0; i < n; i++) {
		fire[i].cb(fire[i].data);
		pthread_mutex_lock(&g_nv_lock);
		if (g_pins[fire[i].slot].live && g_pins[fire[i].slot].pt == fire[i].pt) g_pins[fire[i].slot].cb_done = 1;
		pthread_mutex_unlock(&g_nv_lock);
	}
	return n;
}

/* ------------------------------------------------------------------ mock ib_core (PeerDirect core) */
/*
 * Thread model (what the real PeerDirect core guarantees, so that what is left to the CLIENT is exercised):
 *   * teardown of an MR (dma_unmap + put_pages) happens exactly once, under the MR's lock -- whoever gets there
 *     first, the invalidate upcall (synchronous mode) or dereg;
 *   * release is called once, after the teardown, by dereg -- or, with sim_ib_set_release_in_invalidate(1), from
 *     INSIDE the invalidate upcall (an ib_core that destroys the MR on invalidation), in which case the client's
 *     free callback returns into a context that ib_core has already released;
 *   * nothing orders a dereg on one thread against the client's free callback on another beyond that lock: the
 *     callback's epilogue (after the upcall) may run concurrently with release.
 */
#define MAX_MRS 256
struct sim_mr {
	int live, invalidated, torn_down, released;
	int registering;         /* reg_mr still running: the MR is not published yet, an invalidation may tear the pin down but
	                          * never destroys (releases) it -- the registering thread does, when it gets back */
	u32 gen;                 /* bumped on every (re)use of the slot: core_context = ticket (gen << 16 | slot), never reused */
	pthread_mutex_t lock;
	void *client_ctx;
	struct sg_table sg;
	int nmap;
	unsigned long page_size;
	struct pci_dev pdev;
};
static const struct peer_memory_client *g_client;
static struct sim_mr g_mrs[MAX_MRS];
static pthread_mutex_t g_ib_lock = PTHREAD_MUTEX_INITIALIZER;   /* slot allocation */
static int g_refuse_registration, g_sync_invalidate = 1, g_invalidate_calls, g_release_in_invalidate;
static int g_reg_handle_token;

SIM_API void sim_ib_set_refuse(int on) { g_refuse_registration = on; }
SIM_API void sim_ib_set_sync_invalidate(int on) { g_sync_invalidate = on; }
SIM_API void sim_ib_set_release_in_invalidate(int on) { g_release_in_invalidate = on; }
SIM_API int sim_ib_invalidate_calls(void) { return __atomic_load_n(&g_invalidate_calls, __ATOMIC_SEQ_CST); }
SIM_API const char *sim_ib_client_name(void) { return g_client ? g_client->name : ""; }
SIM_API const char *sim_ib_client_version(void) { return g_client ? g_client->version : ""; }

/* caller holds mr->lock */
static void mr_teardown_locked(struct sim_mr *mr)
{
	if (mr->torn_down) return;
	mr->torn_down = 1;
	g_client->dma_unmap(&mr->sg, mr->client_ctx, &mr->pdev.dev);
	g_client->put_pages(&mr->sg, mr->client_ctx);
}
static int sim_invalidate(void *reg_handle, u64 core_context)
{
	/* core_context is a ticket, as in the real PeerDirect core: a late invalidation of an MR that is already gone
	 * (its slot possibly re-used by another registration) must find nothing */
	const u32 slot = (u32)(core_context & 0xffff), gen = (u32)(core_context >> 16);
	struct sim_mr *mr = slot < MAX_MRS ? &g_mrs[slot] : NULL;
	void *ctx = NULL;

	__atomic_add_fetch(&g_invalidate_calls, 1, __ATOMIC_SEQ_CST);
	if (reg_handle != &g_reg_handle_token || !mr) return -EINVAL;
	pthread_mutex_lock(&mr->lock);
	if (!mr->live || mr->gen != gen) { pthread_mutex_unlock(&mr->lock); return -EINVAL; }
	mr->invalidated = 1;
	if (g_sync_invalidate) mr_teardown_locked(mr); /* re-enters the client from inside its free callback */
	if (g_sync_invalidate && g_release_in_invalidate && !mr->released && !mr->registering) { mr->released = 1; ctx = mr->client_ctx; }
	pthread_mutex_unlock(&mr->lock);
	if (ctx) g_client->release(ctx);               /* the MR is destroyed from inside the upcall */
	return 0;
}
void *ib_register_peer_memory_client(const struct peer_memory_client *c, invalidate_peer_memory *cb)
{
	if (g_refuse_registration || !c || !c->acquire || !c->get_pages || !c->dma_map || !c->dma_unmap || !c->put_pages ||
	    !c->get_page_size || !c->release)
		return NULL;
	g_client = c;
	if (cb) *cb = sim_invalidate;
	return &g_reg_handle_token;
}
void ib_unregister_peer_memory_client(void *h) { if (h == &g_reg_handle_token) g_client = NULL; }

/* ibv_reg_mr on [addr, addr+size) for HCA `dev_id`: ib_core's callback order. */
SIM_API long sim_ib_reg_mr(u64 addr, u64 size, int dev_id)
{
	struct sim_mr *mr = NULL;
	void *ctx = NULL;
	int i, rc;

	if (!g_client) return -ENODEV;
	if (!g_client->acquire((unsigned long)addr, (size_t)size, NULL, NULL, &ctx)) return -EOPNOTSUPP; /* not ours */
	pthread_mutex_lock(&g_ib_lock);
	for (i = 0; i < MAX_MRS; i++) {
		int free_slot;

		pthread_mutex_lock(&g_mrs[i].lock); free_slot = !g_mrs[i].live; pthread_mutex_unlock(&g_mrs[i].lock);
		if (free_slot) { mr = &g_mrs[i]; break; }
	}
	if (mr) {
		/* fields are re-initialised under the slot's own lock: a late invalidation may be looking at it */
		pthread_mutex_lock(&mr->lock);
		mr->gen++;
		mr->invalidated = mr->torn_down = mr->released = 0;
		mr->client_ctx = NULL; mr->nmap = 0; memset(&mr->sg, 0, sizeof mr->sg);
		mr->live = 1; mr->registering = 1;
		pthread_mutex_unlock(&mr->lock);
	}
	pthread_mutex_unlock(&g_ib_lock);
	if (!mr) { g_client->release(ctx); return -ENOMEM; }
	pthread_mutex_lock(&mr->lock);
	mr->client_ctx = ctx;
	mr->pdev.dev.id = dev_id; mr->pdev.vendor = 0x15b3;
	mr->page_size = g_client->get_page_size(mr->client_ctx);
	pthread_mutex_unlock(&mr->lock);
	/* get_pages / dma_map run WITHOUT the MR lock: a revoke may fire (and call sim_invalidate) while they do */
	rc = g_client->get_pages((unsigned long)addr, (size_t)size, 1, 1, &mr->sg, ctx, ((u64)mr->gen << 16) | (u64)i);
	if (!rc) {
		int pinned_only = 0;

		rc = g_client->dma_map(&mr->sg, ctx, &mr->pdev.dev, 0, &mr->nmap);
		pthread_mutex_lock(&mr->lock);
		/* an invalidation that ran meanwhile may have torn the MR down, or even released the context: neither may be
		 * repeated here */
		if (rc && !mr->torn_down && !mr->released) { mr->torn_down = 1; pinned_only = 1; }
		pthread_mutex_unlock(&mr->lock);
		if (pinned_only) g_client->put_pages(&mr->sg, ctx);
	}
	{
		int rel, gone;

		pthread_mutex_lock(&mr->lock);
		mr->registering = 0;
		/* an invalidation arrived while the MR was being registered and this ib_core destroys MRs on invalidation:
		 * it could not do so under our feet, so it falls to us now */
		if (!rc && mr->invalidated && g_release_in_invalidate && !mr->released) { mr_teardown_locked(mr); rc = -EFAULT; }
		gone = mr->released;
		rel = rc && !mr->released;
		if (rel) { mr->released = 1; mr->torn_down = 1; }
		pthread_mutex_unlock(&mr->lock);
		if (rel) g_client->release(ctx);
		if (rc || gone) {
			pthread_mutex_lock(&mr->lock); mr->live = 0; pthread_mutex_unlock(&mr->lock);
			return rc ? rc : -EFAULT;
		}
	}
	return i;
}
SIM_API int sim_ib_mr_nmap(long id) { return g_mrs[id].nmap; }
SIM_API u64 sim_ib_mr_page_size(long id) { return g_mrs[id].page_size; }
SIM_API int sim_ib_mr_invalidated(long id)
{
	int v;

	pthread_mutex_lock(&g_mrs[id].lock); v = g_mrs[id].invalidated; pthread_mutex_unlock(&g_mrs[id].lock);
	return v;
}
SIM_API int sim_ib_mr_dma(long id, int i, u64 *addr, u64 *len)
{
	struct sim_mr *mr = &g_mrs[id];
	int rc = -EINVAL;

	pthread_mutex_lock(&mr->lock);
	if (mr->live && !mr->torn_down && i >= 0 && i < (int)mr->sg.nents) {
		*addr = mr->sg.sgl[i].dma_address; *len = mr->sg.sgl[i].dma_length;
		rc = 0;
	}
	pthread_mutex_unlock(&mr->lock);
	return rc;
}
SIM_API int sim_ib_dereg_mr(long id)
{
	struct sim_mr *mr = &g_mrs[id];
	void *ctx = NULL;

	if (id < 0 || id >= MAX_MRS) return -EINVAL;
	pthread_mutex_lock(&mr->lock);
	if (!mr->live) { pthread_mutex_unlock(&mr->lock); return -EINVAL; }
	mr_teardown_locked(mr);
	if (!mr->released) { mr->released = 1; ctx = mr->client_ctx; }
	pthread_mutex_unlock(&mr->lock);
	if (ctx) g_client->release(ctx);
	pthread_mutex_lock(&mr->lock); mr->live = 0; pthread_mutex_unlock(&mr->lock);
	return 0;
}
/* Out-of-order / malformed sequences the real ib_core never produces but the client must survive. */
SIM_API int sim_ib_bad_sequence(int which, u64 addr, u64 size)
{
	void *ctx = NULL;
	struct sg_table sg = {0};
	struct pci_dev pdev = { { 7, "hca" }, 0x15b3, 0 };
	int nmap = 0, rc = 0;

	if (!g_client) return -ENODEV;
	if (!g_client->acquire((unsigned long)addr, (size_t)size, NULL, NULL, &ctx)) return -EOPNOTSUPP;
	switch (which) {
	case 0: rc = g_client->dma_map(&sg, ctx, &pdev.dev, 0, &nmap); break;                        /* dma_map before get_pages */
	case 1: rc = g_client->get_pages((unsigned long)addr + 65536, (size_t)size, 1, 1, &sg, ctx, 1); break; /* address mismatch */
	case 2: rc = g_client->get_pages((unsigned long)addr, (size_t)size + 65536, 1, 1, &sg, ctx, 1); break; /* size mismatch */
	case 3: /* double get_pages */
		rc = g_client->get_pages((unsigned long)addr, (size_t)size, 1, 1, &sg, ctx, 1);
		if (!rc) rc = g_client->get_pages((unsigned long)addr, (size_t)size, 1, 1, &sg, ctx, 1);
		g_client->put_pages(&sg, ctx);
		break;
	case 4: /* release with the pin and mapping still in place */
		rc = g_client->get_pages((unsigned long)addr, (size_t)size, 1, 1, &sg, ctx, 1);
		if (!rc) rc = g_client->dma_map(&sg, ctx, &pdev.dev, 0, &nmap);
		g_client->release(ctx);
		if (sg.sgl) sg_free_table(&sg);
		return rc;
	case 5: rc = g_client->dma_map(&sg, ctx, NULL, 0, &nmap); break;                             /* no DMA device */
	}
	g_client->release(ctx);
	return rc;
}

/* ------------------------------------------------------------------ misc device + file glue */
static struct miscdevice *g_misc;
static int g_misc_fail;
int misc_register(struct miscdevice *m) { if (g_misc_fail) return -EBUSY; g_misc = m; return 0; }
void misc_deregister(struct miscdevice *m) { if (g_misc == m) g_misc = NULL; }
SIM_API void sim_misc_set_fail(int on) { g_misc_fail = on; }
SIM_API const char *sim_dev_name(void) { return g_misc ? g_misc->name : ""; }
SIM_API int sim_dev_mode(void) { return g_misc ? g_misc->mode : -1; }

SIM_API void *sim_dev_open(void)
{
	struct file *f;
	struct inode ino = {0};

	if (!g_misc) return NULL;
	f = calloc(1, sizeof(*f));
	if (g_misc->fops->open(&ino, f)) { free(f); return NULL; }
	return f;
}
SIM_API long sim_dev_ioctl(void *file, unsigned int cmd, void *arg)
{
	if (!g_misc || !file) return -ENODEV;
	return g_misc->fops->unlocked_ioctl(file, cmd, (unsigned long)arg);
}
SIM_API int sim_dev_close(void *file)
{
	struct inode ino = {0};
	int rc;

	if (!file) return -EINVAL;
	rc = g_misc ? g_misc->fops->release(&ino, file) : -ENODEV;
	free(file);
	return rc;
}
struct remap_log { int n, cap; u64 user[4096]; u64 pfn[4096]; u64 size[4096]; };
int remap_pfn_range(struct vm_area_struct *vma, unsigned long addr, unsigned long pfn, unsigned long size, pgprot_t prot)
{
	struct remap_log *l = vma->sim_log;

	if (!l || l->n >= 4096) return -ENOMEM;
	l->user[l->n] = addr; l->pfn[l->n] = pfn; l->size[l->n] = size; l->n++;
	return 0;
}
/* mmap(fd, len, offset = gpu_va): returns the number of remap calls; out[] gets (user_off, bus, size) triples. */
SIM_API int sim_dev_mmap(void *file, u64 gpu_va, u64 len, u64 *out, int max_triples)
{
	struct remap_log *l = calloc(1, sizeof(*l));
	struct vm_area_struct vma = {0};
	int rc, i;

	if (!g_misc || !file) { free(l); return -ENODEV; }
	vma.vm_start = 0x7f0000000000UL; vma.vm_end = vma.vm_start + len; vma.vm_pgoff = gpu_va >> PAGE_SHIFT; vma.sim_log = l;
	rc = g_misc->fops->mmap(file, &vma);
	if (rc) { free(l); return rc; }
	for (i = 0; i < l->n && i < max_triples; i++) {
		out[3 * i] = l->user[i] - vma.vm_start; out[3 * i + 1] = l->pfn[i] << PAGE_SHIFT; out[3 * i + 2] = l->size[i];
	}
	rc = l->n;
	free(l);
	return rc;
}

/* ------------------------------------------------------------------ module entry points */
int sim_init_b200p2p(void);
void sim_exit_b200p2p(void);
int sim_init_b200p2ptest(void);
void sim_exit_b200p2ptest(void);
SIM_API int sim_b200p2p_load(void) { return sim_init_b200p2p(); }
SIM_API void sim_b200p2p_unload(void) { sim_exit_b200p2p(); }
SIM_API int sim_b200p2ptest_load(void) { return sim_init_b200p2ptest(); }
SIM_API void sim_b200p2ptest_unload(void) { sim_exit_b200p2ptest(); }

SIM_API void sim_reset(void)
{
	memset(g_allocs, 0, sizeof g_allocs);
	memset(g_pins, 0, sizeof g_pins);
	{
		static int locks_ready;
		int k;

		for (k = 0; k < MAX_MRS; k++) {
			pthread_mutex_t keep = g_mrs[k].lock;
			memset(&g_mrs[k], 0, sizeof g_mrs[k]);
			if (locks_ready) g_mrs[k].lock = keep; else pthread_mutex_init(&g_mrs[k].lock, NULL);
		}
		locks_ready = 1;
	}
	memset(g_dbgfs, 0, sizeof g_dbgfs);
	g_release_in_invalidate = 0; g_dma_map_hook = NULL; g_concurrent = 0;
	g_live_pt = g_live_map = g_misuse = 0;
	g_fail_next_get_pages = g_fail_next_dma_map = g_revoke_during_get_pages = 0;
	g_refuse_registration = 0; g_sync_invalidate = 1; g_invalidate_calls = 0;
	g_copy_fault_in = -1; g_lock_errors = 0; g_misc_fail = 0;
	sim_log_reset();
}
