// SPDX-License-Identifier: MIT
/*
 * Multi-threaded life-cycle stress of both kernel modules under the userspace simulation (kmod/shim), built to run
 * under ThreadSanitizer and AddressSanitizer (make -C kmod check-tsan check-stress).  The single-threaded scenarios of
 * tests/test_kmod_sim.py cannot see the races that matter in a peer-memory client: the GPU driver's free callback
 * runs on its own thread while ib_core registers, deregisters and releases on others.
 *
 *   registrars   N threads: ibv_reg_mr (acquire -> get_page_size -> get_pages -> dma_map) on a random sub-range of a
 *                random "GPU allocation", hold it a moment, ibv_dereg_mr (dma_unmap -> put_pages -> release)
 *   gpu driver   1 thread: cudaFree of a random allocation -- every pin on it is revoked: free callback ->
 *                invalidate upcall -> (synchronous mode) ib_core tears the MR down from inside the callback, and in
 *                "release inside invalidate" mode even releases the client's context there -- then cudaMalloc again
 *   harness      M threads on /dev/b200p2ptest: open, GET_PAGES x k, GET_BUS_ADDRS, mmap, PUT_PAGES / close with leaks
 *
 * Pass criteria (besides TSan / ASan staying silent): no misuse of the NVIDIA interface as the mock defines it (put
 * or unmap after a COMPLETED revoke, free of an unknown page table, a freed page table handed to the driver), no page
 * table / DMA mapping / kernel allocation left behind, module reference count back to zero, no lock errors.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../include/b200p2ptest.h"

typedef uint64_t u64;
int sim_b200p2p_load(void); void sim_b200p2p_unload(void);
int sim_b200p2ptest_load(void); void sim_b200p2ptest_unload(void);
void sim_reset(void);
int sim_gpu_alloc(u64 va, u64 size); int sim_gpu_free(u64 va);
long sim_ib_reg_mr(u64 addr, u64 size, int dev_id); int sim_ib_dereg_mr(long id);
int sim_ib_mr_dma(long id, int i, u64 *addr, u64 *len);
void sim_ib_set_sync_invalidate(int on); void sim_ib_set_release_in_invalidate(int on); void sim_nv_set_concurrent(int on);
int sim_live_page_tables(void); int sim_live_dma_mappings(void); int sim_nv_misuse(void); int sim_lock_errors(void);
int sim_module_refcount(void); long sim_live_allocs(void); int sim_live_pins(void); int sim_ib_invalidate_calls(void);
void *sim_dev_open(void); long sim_dev_ioctl(void *f, unsigned int cmd, void *arg); int sim_dev_close(void *f);
int sim_dev_mmap(void *f, u64 gpu_va, u64 len, u64 *out, int max_triples);
int sim_debugfs_read(const char *name, char *out, int cap);

#define PAGE 65536ull
#define N_ALLOC 6
#define ALLOC_PAGES 16
static const u64 BASE = 0x7f0000000000ull;
static int g_stop;
#define STOPPED() __atomic_load_n(&g_stop, __ATOMIC_ACQUIRE)
static unsigned long g_regs, g_reg_fail, g_frees, g_ioctls;

static u64 alloc_va(int k) { return BASE + (u64)k * 64 * PAGE; }
static unsigned rnd(unsigned *s) { *s = *s * 1103515245u + 12345u; return *s >> 8; }
static void nap(unsigned *s) { struct timespec ts = {0, (long)(rnd(s) % 20000)}; nanosleep(&ts, NULL); }

static void *registrar(void *arg)
{
	unsigned seed = (unsigned)(uintptr_t)arg * 7919u + 17u;

	while (!STOPPED()) {
		int k = (int)(rnd(&seed) % N_ALLOC);
		u64 off = (rnd(&seed) % (ALLOC_PAGES - 4)) * PAGE + (rnd(&seed) % 3) * 4096;
		u64 len = (1 + rnd(&seed) % 3) * PAGE + (rnd(&seed) % 2) * 100;
		long mr = sim_ib_reg_mr(alloc_va(k) + off, len, (int)(rnd(&seed) % 4));

		if (mr < 0) { __atomic_add_fetch(&g_reg_fail, 1, __ATOMIC_RELAXED); continue; }   /* freed under us: fine */
		__atomic_add_fetch(&g_regs, 1, __ATOMIC_RELAXED);
		u64 a, l;
		sim_ib_mr_dma(mr, 0, &a, &l);
		if (rnd(&seed) & 1) nap(&seed);
		sim_ib_dereg_mr(mr);
	}
	return NULL;
}
static void *gpu_driver(void *arg)
{
	unsigned seed = 4242;

	(void)arg;
	while (!STOPPED()) {
		int k = (int)(rnd(&seed) % N_ALLOC);

		if (sim_gpu_free(alloc_va(k)) >= 0) __atomic_add_fetch(&g_frees, 1, __ATOMIC_RELAXED);
		nap(&seed);
		sim_gpu_alloc(alloc_va(k), ALLOC_PAGES * PAGE);
		nap(&seed);
	}
	return NULL;
}
static void *harness_user(void *arg)
{
	unsigned seed = (unsigned)(uintptr_t)arg * 104729u + 3u;
	u64 out[30];

	while (!STOPPED()) {
		void *f = sim_dev_open();
		int n = 1 + (int)(rnd(&seed) % 4), i;

		if (!f) continue;
		for (i = 0; i < n; i++) {
			int k = (int)(rnd(&seed) % N_ALLOC);
			struct b200p2p_get_pages g = { .addr = alloc_va(k) + (rnd(&seed) % 8) * PAGE, .length = (1 + rnd(&seed) % 4) * PAGE };

			__atomic_add_fetch(&g_ioctls, 1, __ATOMIC_RELAXED);
			if (sim_dev_ioctl(f, B200P2PTEST_IOCTL_GET_PAGES, &g)) continue;
			struct b200p2p_get_bus_addrs *b = calloc(1, sizeof(*b));
			b->handle = g.handle; b->count = 4;
			sim_dev_ioctl(f, B200P2PTEST_IOCTL_GET_BUS_ADDRS, b);
			free(b);
			sim_dev_mmap(f, g.addr, PAGE, out, 10);
			if (rnd(&seed) & 1) {
				struct b200p2p_put_pages p = { .addr = g.addr, .length = g.length };
				sim_dev_ioctl(f, B200P2PTEST_IOCTL_PUT_PAGES, &p);
			}
		}
		sim_dev_close(f);       /* with whatever pins were left: release must unpin them */
	}
	return NULL;
}

int main(int argc, char **argv)
{
	double seconds = argc > 1 ? atof(argv[1]) : 2.0;
	extern void sim_set_verbose(int v);
	if (getenv("SIM_VERBOSE")) sim_set_verbose(1);
	int mode, fails = 0;

	/* mode 0: deferred teardown; 1: synchronous teardown from the upcall; 2: + release inside the upcall */
	for (mode = 0; mode < 3; mode++) {
		pthread_t th[16];
		int n = 0, i;
		long base;
		char stats[512];

		sim_reset();
		base = sim_live_allocs();
		__atomic_store_n(&g_stop, 0, __ATOMIC_RELEASE); g_regs = g_reg_fail = g_frees = g_ioctls = 0;
		sim_ib_set_sync_invalidate(mode >= 1);
		sim_ib_set_release_in_invalidate(mode == 2);
		sim_nv_set_concurrent(1);
		if (sim_b200p2p_load() || sim_b200p2ptest_load()) { fprintf(stderr, "module load failed\n"); return 2; }
		for (i = 0; i < N_ALLOC; i++) sim_gpu_alloc(alloc_va(i), ALLOC_PAGES * PAGE);
		for (i = 0; i < 4; i++) pthread_create(&th[n++], NULL, registrar, (void *)(uintptr_t)(i + 1));
		for (i = 0; i < 3; i++) pthread_create(&th[n++], NULL, harness_user, (void *)(uintptr_t)(i + 1));
		pthread_create(&th[n++], NULL, gpu_driver, NULL);
		usleep((useconds_t)(seconds * 1e6));
		__atomic_store_n(&g_stop, 1, __ATOMIC_RELEASE);
		for (i = 0; i < n; i++) pthread_join(th[i], NULL);
		for (i = 0; i < N_ALLOC; i++) sim_gpu_free(alloc_va(i));
		stats[0] = 0;
		sim_debugfs_read("stats", stats, sizeof stats);
		for (char *c = stats; *c; ++c) if (*c == '\n') *c = ' ';
		sim_b200p2ptest_unload();
		sim_b200p2p_unload();
		printf("mode %d: %lu registrations (%lu refused mid-free), %lu frees, %d invalidations, %lu harness pins | %s\n", mode, g_regs,
		       g_reg_fail, g_frees, sim_ib_invalidate_calls(), g_ioctls, stats);
		if (sim_live_page_tables() || sim_live_dma_mappings() || sim_nv_misuse() || sim_lock_errors() || sim_module_refcount() ||
		    sim_live_allocs() != base || sim_live_pins()) {
			printf("  FAIL: page tables %d, dma mappings %d, misuse %d, lock errors %d, module refs %d, allocations %+ld, pins %d\n",
			       sim_live_page_tables(), sim_live_dma_mappings(), sim_nv_misuse(), sim_lock_errors(), sim_module_refcount(),
			       sim_live_allocs() - base, sim_live_pins());
			fails++;
		}
		if (!g_regs || !g_frees) { printf("  FAIL: the stress did not exercise anything\n"); fails++; }
	}
	printf(fails ? "STRESS FAILED\n" : "STRESS OK\n");
	return fails ? 1 : 0;
}
