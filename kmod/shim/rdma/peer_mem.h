/* SPDX-License-Identifier: MIT */
/*
 * Minimal declarations of the MLNX_OFED PeerDirect interface (<rdma/peer_mem.h>), reconstructed from
 * how peer-memory clients use it (the reference's call sites: amdp2p.c:363-371, :388-391, :103, :407).
 * A real build picks up OFED's header (kmod/Makefile: OFA_KERNEL_DIR); this copy exists so the module
 * still compile-checks, and so the userspace simulation can stand in for ib_core.
 */
#ifndef B200_SHIM_PEER_MEM_H_
#define B200_SHIM_PEER_MEM_H_

#include <linux/types.h>
#include <linux/scatterlist.h>

struct device;

#define IB_PEER_MEMORY_NAME_MAX 64
#define IB_PEER_MEMORY_VER_MAX 16

struct peer_memory_client {
	char name[IB_PEER_MEMORY_NAME_MAX];
	char version[IB_PEER_MEMORY_VER_MAX];
	int (*acquire)(unsigned long addr, size_t size, void *peer_mem_private_data, char *peer_mem_name,
		       void **client_context);
	int (*get_pages)(unsigned long addr, size_t size, int write, int force, struct sg_table *sg_head,
			 void *client_context, u64 core_context);
	int (*dma_map)(struct sg_table *sg_head, void *client_context, struct device *dma_device, int dmasync,
		       int *nmap);
	int (*dma_unmap)(struct sg_table *sg_head, void *client_context, struct device *dma_device);
	void (*put_pages)(struct sg_table *sg_head, void *client_context);
	unsigned long (*get_page_size)(void *client_context);
	void (*release)(void *client_context);
	void *(*get_context_private_data)(u64 peer_id);
	void (*put_context_private_data)(void *context);
};

typedef int (*invalidate_peer_memory)(void *reg_handle, u64 core_context);

void *ib_register_peer_memory_client(const struct peer_memory_client *peer_client,
				     invalidate_peer_memory *invalidate_callback);
void ib_unregister_peer_memory_client(void *reg_handle);

#endif
