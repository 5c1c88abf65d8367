/* SPDX-License-Identifier: MIT */
/*
 * b200p2ptest - user <-> kernel ABI of the B200 GPU P2P test harness.
 *
 * Same four verbs as the reference harness (include/amdp2ptest.h:33-72): is-GPU-address,
 * page-size query, pin, unpin -- plus a read-back of the bus addresses, and an mmap() whose
 * page offset is the GPU virtual address (tests/amdp2ptest.c:336-395).
 *
 * Deliberately NOT carried over from the reference:
 *   - ioctl numbers that encode sizeof(pointer) instead of sizeof(struct)
 *   - an IS_GPU_ADDRESS request declared _IOW although the kernel writes the answer back
 *   - a request naming a struct that does not exist, and the misspelt AMD2P2PTEST_ prefix
 * Every request below is _IOWR(magic, nr, struct) with the struct it really carries.
 * All fields are fixed-width; the layout is identical for 32- and 64-bit callers.
 */
#ifndef B200P2PTEST_H_
#define B200P2PTEST_H_

#ifdef __KERNEL__
#include <linux/ioctl.h>
#include <linux/types.h>
#else
#include <stdint.h>
#include <sys/ioctl.h>
#endif

#define B200P2PTEST_IOCTL_MAGIC 'B'
#define B200P2PTEST_DEVICE_NAME "b200p2ptest"
#define B200P2PTEST_DEVICE_PATH "/dev/b200p2ptest"
#define B200P2PTEST_ABI_VERSION 1

/* NVIDIA's P2P interface pins in units of 64 KiB GPU pages. */
#define B200P2P_GPU_PAGE_SHIFT 16
#define B200P2P_GPU_PAGE_SIZE (1ULL << B200P2P_GPU_PAGE_SHIFT)
#define B200P2P_MAX_BUS_ADDRS 512

struct b200p2p_is_gpu_address {
	uint64_t addr;      /* in  */
	uint32_t ret_value; /* out: 1 = GPU virtual address of the calling process */
	uint32_t reserved;
};

struct b200p2p_get_page_size {
	uint64_t addr;      /* in  */
	uint64_t length;    /* in  */
	uint64_t page_size; /* out: bytes */
};

struct b200p2p_get_pages {
	uint64_t addr;      /* in: must be 64 KiB aligned */
	uint64_t length;    /* in: multiple of 64 KiB */
	uint64_t handle;    /* out: opaque id of this pin (the same range may be pinned repeatedly) */
	uint32_t entries;   /* out: number of GPU pages pinned */
	uint32_t page_size; /* out: bytes per entry */
};

struct b200p2p_put_pages {
	uint64_t addr;      /* in: every pin with exactly this addr+length is released */
	uint64_t length;    /* in  */
	uint32_t released;  /* out: how many pins matched */
	uint32_t reserved;
};

struct b200p2p_get_bus_addrs {
	uint64_t handle;    /* in: from GET_PAGES */
	uint32_t first;     /* in: first entry wanted */
	uint32_t count;     /* in: entries wanted (<= B200P2P_MAX_BUS_ADDRS); out: entries returned */
	uint64_t addrs[B200P2P_MAX_BUS_ADDRS]; /* out: bus (DMA) address of each GPU page */
};

#define B200P2PTEST_IOCTL_IS_GPU_ADDRESS _IOWR(B200P2PTEST_IOCTL_MAGIC, 1, struct b200p2p_is_gpu_address)
#define B200P2PTEST_IOCTL_GET_PAGE_SIZE  _IOWR(B200P2PTEST_IOCTL_MAGIC, 2, struct b200p2p_get_page_size)
#define B200P2PTEST_IOCTL_GET_PAGES      _IOWR(B200P2PTEST_IOCTL_MAGIC, 3, struct b200p2p_get_pages)
#define B200P2PTEST_IOCTL_PUT_PAGES      _IOWR(B200P2PTEST_IOCTL_MAGIC, 4, struct b200p2p_put_pages)
#define B200P2PTEST_IOCTL_GET_BUS_ADDRS  _IOWR(B200P2PTEST_IOCTL_MAGIC, 5, struct b200p2p_get_bus_addrs)

#endif /* B200P2PTEST_H_ */
