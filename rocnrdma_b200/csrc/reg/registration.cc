// N1 userspace registration helpers: dma-buf export of GPU HBM and pointer classification through the
// CUDA driver API (entry points resolved at run time, so the library links without libcuda).
//
// The dma-buf route is the modern no-kernel-module way to hand GPU pages to an HCA:
//     cuMemGetHandleForAddressRange(..., CU_MEM_RANGE_HANDLE_TYPE_DMA_BUF_FD) -> fd -> ibv_reg_dmabuf_mr(pd, 0, len, iova, fd, access)
// It plays the role of get_pages + dma_map of the reference's bridge (amdp2p.c:169-264) with the pin
// owned by the fd: closing it is put_pages (amdp2p.c:283-313).
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#define RN_API extern "C" __attribute__((visibility("default")))

namespace {
template <typename Fn>
Fn entry(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return (Fn)p;
}
typedef CUresult (*GetHandleForRangeFn)(void*, CUdeviceptr, size_t, CUmemRangeHandleType, unsigned long long);
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
typedef CUresult (*PointerGetAttributeFn)(void*, CUpointer_attribute, CUdeviceptr);
typedef CUresult (*DeviceGetAttributeFn)(int*, CUdevice_attribute, CUdevice);
}  // namespace

// GPU pages are pinned in 64 KiB units by the NVIDIA P2P interface; dma-buf export wants host-page
// (4 KiB) alignment of address and size.
RN_API uint64_t rn_gpu_page_size() { return 65536; }

// Export [ptr, ptr+len) as a dma-buf fd.  Returns the fd (>= 0) or -errno-style negative code:
//   -38 driver lacks the entry point, -95 device lacks dma-buf support, -22 misaligned, -5 driver error.
RN_API int rn_dmabuf_export(uint64_t ptr, uint64_t len, int* cu_err_out) {
  if (cu_err_out) *cu_err_out = 0;
  if (!ptr || !len || (ptr & 4095) || (len & 4095)) return -22;
  static GetHandleForRangeFn fn = entry<GetHandleForRangeFn>("cuMemGetHandleForAddressRange");
  if (!fn) return -38;
  int fd = -1;
  CUresult r = fn(&fd, (CUdeviceptr)ptr, (size_t)len, CU_MEM_RANGE_HANDLE_TYPE_DMA_BUF_FD, 0);
  if (r != CUDA_SUCCESS) {
    if (cu_err_out) *cu_err_out = (int)r;
    return r == CUDA_ERROR_NOT_SUPPORTED ? -95 : -5;
  }
  return fd;
}
RN_API int rn_dmabuf_close(int fd) { return fd >= 0 ? close(fd) : -9; }
// Size the kernel reports for the exported buffer (lseek SEEK_END on a dma-buf fd), or negative.
RN_API int64_t rn_dmabuf_size(int fd) {
  off_t end = lseek(fd, 0, SEEK_END);
  if (end < 0) return -errno;
  lseek(fd, 0, SEEK_SET);
  return (int64_t)end;
}

// The allocation that contains ptr: base and size (for IPC export and for widening to page boundaries).
RN_API int rn_alloc_range(uint64_t ptr, uint64_t* base, uint64_t* size) {
  static GetAddressRangeFn fn = entry<GetAddressRangeFn>("cuMemGetAddressRange");
  if (!fn) return -38;
  CUdeviceptr b = 0;
  size_t s = 0;
  if (fn(&b, &s, (CUdeviceptr)ptr) != CUDA_SUCCESS) return -22;
  *base = (uint64_t)b;
  *size = (uint64_t)s;
  return 0;
}

// Capability flags of a device: bit0 dma-buf, bit1 GPUDirect RDMA, bit2 VMM, bit3 posix-fd handles.
RN_API int rn_device_caps(int dev) {
  static DeviceGetAttributeFn fn = entry<DeviceGetAttributeFn>("cuDeviceGetAttribute");
  if (!fn) return -38;
  int caps = 0, v = 0;
  if (fn(&v, CU_DEVICE_ATTRIBUTE_DMA_BUF_SUPPORTED, dev) == CUDA_SUCCESS && v) caps |= 1;
  if (fn(&v, CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_SUPPORTED, dev) == CUDA_SUCCESS && v) caps |= 2;
  if (fn(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev) == CUDA_SUCCESS && v) caps |= 4;
  if (fn(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) == CUDA_SUCCESS && v) caps |= 8;
  return caps;
}

RN_API int rn_device_pci(int dev, char* out, int n) {
  return cudaDeviceGetPCIBusId(out, n, dev) == cudaSuccess ? 0 : -19;
}

// Identity of the allocation behind a device pointer (CU_POINTER_ATTRIBUTE_BUFFER_ID): unique per allocation for
// the life of the process, so "the memory under this registration was freed" -- and even "freed and the address
// handed out again" -- is observable from userspace.  Returns 0 and *id, or -22 when the driver no longer knows
// the pointer (it was freed).  This is the userspace stand-in for the kernel-side free callback
// (nvidia_p2p_get_pages' free_callback; the reference's free_callback, amdp2p.c:88-109).
RN_API int rn_buffer_id(uint64_t ptr, uint64_t* id) {
  static PointerGetAttributeFn fn = entry<PointerGetAttributeFn>("cuPointerGetAttribute");
  if (!fn) return -38;
  unsigned long long v = 0;
  if (fn(&v, CU_POINTER_ATTRIBUTE_BUFFER_ID, (CUdeviceptr)ptr) != CUDA_SUCCESS) return -22;
  *id = (uint64_t)v;
  return 0;
}
