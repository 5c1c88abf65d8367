/* SPDX-License-Identifier: MIT */
/*
 * Minimal declarations of NVIDIA's GPUDirect RDMA kernel interface (nv-p2p.h), written from the
 * public API description, for compile-checking b200p2p / b200p2ptest where the driver's own
 * header is not installed.  A real build uses the header shipped with the NVIDIA driver
 * (kmod/Makefile prefers it); the structures below only name the fields this code touches.
 */
#ifndef B200_SHIM_NV_P2P_H_
#define B200_SHIM_NV_P2P_H_

#include <linux/types.h>

struct pci_dev;

enum nvidia_p2p_page_size_type {
	NVIDIA_P2P_PAGE_SIZE_4KB = 0,
	NVIDIA_P2P_PAGE_SIZE_64KB,
	NVIDIA_P2P_PAGE_SIZE_128KB,
	NVIDIA_P2P_PAGE_SIZE_COUNT
};

struct nvidia_p2p_page {
	uint64_t physical_address;
};

struct nvidia_p2p_page_table {
	uint32_t version;
	uint32_t page_size; /* enum nvidia_p2p_page_size_type */
	struct nvidia_p2p_page **pages;
	uint32_t entries;
	uint8_t *gpu_uuid;
};

struct nvidia_p2p_dma_mapping {
	uint32_t version;
	enum nvidia_p2p_page_size_type page_size_type;
	uint32_t entries;
	uint64_t *dma_addresses;
	void *private;
	struct pci_dev *pci_dev;
};

int nvidia_p2p_get_pages(uint64_t p2p_token, uint32_t va_space, uint64_t virtual_address, uint64_t length,
			 struct nvidia_p2p_page_table **page_table, void (*free_callback)(void *data), void *data);
int nvidia_p2p_put_pages(uint64_t p2p_token, uint32_t va_space, uint64_t virtual_address,
			 struct nvidia_p2p_page_table *page_table);
int nvidia_p2p_free_page_table(struct nvidia_p2p_page_table *page_table);
int nvidia_p2p_dma_map_pages(struct pci_dev *peer, struct nvidia_p2p_page_table *page_table,
			     struct nvidia_p2p_dma_mapping **dma_mapping);
int nvidia_p2p_dma_unmap_pages(struct pci_dev *peer, struct nvidia_p2p_page_table *page_table,
			       struct nvidia_p2p_dma_mapping *dma_mapping);
int nvidia_p2p_free_dma_mapping(struct nvidia_p2p_dma_mapping *dma_mapping);

#endif
