// Userspace twin of the b200p2ptest harness: the same four verbs (is-GPU-address, page size, pin,
// unpin) plus the CPU window, implemented on the CUDA driver API so the test matrix of SURVEY.md
// section 4.2 runs on a box where no module can be loaded (this one: no /lib/modules, no CAP_SYS_MODULE).
//
//   kernel harness (kmod/b200p2ptest.c)           twin
//   nvidia_p2p_get_pages(va, len)                  dma-buf export of the range (the pin lives in the fd)
//   nvidia_p2p_put_pages                           close(fd)
//   probe pin for is_gpu_address                   cudaPointerGetAttributes
//   mmap of bus addresses                          mmap() of the pin's dma-buf fd where the exporter offers it (a real CPU
//                                                  window through the BAR: rn_p2p_mmap); otherwise peek/poke through
//                                                  cudaMemcpy -- which proves nothing about the aperture, and says so
//                                                  (rn_p2p_window_kind)
//   per-fd list, release-on-close                  per-session list, rn_p2p_close releases leftovers
// The reference ships only the kernel half (tests/amdp2ptest.c) and no program to drive it.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <mutex>
#include <vector>

#define RN_API extern "C" __attribute__((visibility("default")))

extern "C" int rn_dmabuf_export(uint64_t ptr, uint64_t len, int* cu_err_out);
extern "C" int rn_dmabuf_close(int fd);
extern "C" int64_t rn_dmabuf_size(int fd);

namespace {
constexpr uint64_t kGpuPage = 65536;
struct Pin { uint64_t handle, va, size; int fd; void* cpu = nullptr; int map_errno = 0; };
struct Session {
  std::mutex mu;
  std::vector<Pin> pins;
  uint64_t next_handle = 1;
};
bool is_device_ptr(uint64_t p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, (void*)p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice;
}
}  // namespace

RN_API void* rn_p2p_open() { return new Session(); }

RN_API int rn_p2p_close(void* s_) {
  Session* s = (Session*)s_;
  if (!s) return -22;
  int n = 0;
  {
    std::lock_guard<std::mutex> g(s->mu);
    for (auto& p : s->pins) { if (p.cpu) munmap(p.cpu, p.size); rn_dmabuf_close(p.fd); ++n; }
    s->pins.clear();
  }
  delete s;
  return n;   // pins the application left behind (released here, as the kernel harness does on close)
}

RN_API int rn_p2p_is_gpu_address(void* s, uint64_t addr) { (void)s; return is_device_ptr(addr) ? 1 : 0; }

RN_API int rn_p2p_get_page_size(void* s, uint64_t addr, uint64_t len, uint64_t* out) {
  (void)s;
  if (!len || !is_device_ptr(addr) || !is_device_ptr(addr + len - 1)) return -EFAULT;
  *out = kGpuPage;
  return 0;
}

RN_API int rn_p2p_get_pages(void* s_, uint64_t addr, uint64_t len, uint64_t* handle, uint32_t* entries, uint32_t* page_size) {
  Session* s = (Session*)s_;
  if (!len || (addr & (kGpuPage - 1)) || (len & (kGpuPage - 1))) return -EINVAL;
  if (!is_device_ptr(addr) || !is_device_ptr(addr + len - 1)) return -EFAULT;
  int cu = 0;
  int fd = rn_dmabuf_export(addr, len, &cu);
  if (fd < 0) return fd == -95 ? -EOPNOTSUPP : -EFAULT;
  std::lock_guard<std::mutex> g(s->mu);
  Pin p{s->next_handle++, addr, len, fd};
  s->pins.push_back(p);
  *handle = p.handle;
  *entries = (uint32_t)(len / kGpuPage);
  *page_size = (uint32_t)kGpuPage;
  return 0;
}

// Every pin with exactly this addr+len is released (the reference's "same memory pinned several times" rule).
RN_API int rn_p2p_put_pages(void* s_, uint64_t addr, uint64_t len) {
  Session* s = (Session*)s_;
  std::lock_guard<std::mutex> g(s->mu);
  int n = 0;
  for (size_t i = 0; i < s->pins.size();) {
    if (s->pins[i].va == addr && s->pins[i].size == len) {
      if (s->pins[i].cpu) munmap(s->pins[i].cpu, s->pins[i].size);
      rn_dmabuf_close(s->pins[i].fd);
      s->pins.erase(s->pins.begin() + i);
      ++n;
    } else {
      ++i;
    }
  }
  return n;
}

RN_API int rn_p2p_live_pins(void* s_) {
  Session* s = (Session*)s_;
  std::lock_guard<std::mutex> g(s->mu);
  return (int)s->pins.size();
}

// What the kernel says the pinned object is: its size through the dma-buf fd.
RN_API int64_t rn_p2p_pin_size(void* s_, uint64_t handle) {
  Session* s = (Session*)s_;
  std::lock_guard<std::mutex> g(s->mu);
  for (auto& p : s->pins)
    if (p.handle == handle) return rn_dmabuf_size(p.fd);
  return -ENOENT;
}

// The harness's mmap: a CPU mapping of the pinned pages themselves (kmod/b200p2ptest.c remaps the bus addresses; the
// reference: tests/amdp2ptest.c:336-395).  From userspace the only handle on the pin is its dma-buf, so this is
// mmap(fd): it works exactly when the exporting driver implements the dma-buf mmap op.  Returns 0 and the address, or
// -errno of the attempt (remembered per pin: rn_p2p_window_kind reports which window peek / poke are using).
RN_API int rn_p2p_mmap(void* s_, uint64_t handle, uint64_t* cpu_addr, uint64_t* len) {
  Session* s = (Session*)s_;
  std::lock_guard<std::mutex> g(s->mu);
  for (auto& p : s->pins) {
    if (p.handle != handle) continue;
    if (!p.cpu && !p.map_errno) {
      void* m = mmap(nullptr, p.size, PROT_READ | PROT_WRITE, MAP_SHARED, p.fd, 0);
      if (m == MAP_FAILED) p.map_errno = errno ? errno : EIO;
      else p.cpu = m;
    }
    if (!p.cpu) return -p.map_errno;
    *cpu_addr = (uint64_t)p.cpu;
    *len = p.size;
    return 0;
  }
  return -ENOENT;
}
// 1: peek / poke at gpu_va go through a CPU mapping of the pin (the aperture is live); 0: through cudaMemcpy
RN_API int rn_p2p_window_kind(void* s_, uint64_t gpu_va) {
  Session* s = (Session*)s_;
  std::lock_guard<std::mutex> g(s->mu);
  for (auto& p : s->pins)
    if (gpu_va >= p.va && gpu_va < p.va + p.size) return p.cpu ? 1 : 0;
  return -EINVAL;
}

// CPU window: read / write `n` bytes at gpu_va, which must lie inside one live pin.
static int window(Session* s, uint64_t gpu_va, uint64_t n, uint8_t** cpu) {
  std::lock_guard<std::mutex> g(s->mu);
  for (auto& p : s->pins)
    if (gpu_va >= p.va && gpu_va + n <= p.va + p.size) {
      *cpu = p.cpu ? (uint8_t*)p.cpu + (gpu_va - p.va) : nullptr;
      return 0;
    }
  return -EINVAL;
}
RN_API int rn_p2p_peek(void* s, uint64_t gpu_va, void* out, uint64_t n) {
  uint8_t* cpu = nullptr;
  int rc = window((Session*)s, gpu_va, n, &cpu);
  if (rc) return rc;
  if (cpu) { memcpy(out, cpu, n); return 0; }
  return cudaMemcpy(out, (const void*)gpu_va, n, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -EIO;
}
RN_API int rn_p2p_poke(void* s, uint64_t gpu_va, const void* in, uint64_t n) {
  uint8_t* cpu = nullptr;
  int rc = window((Session*)s, gpu_va, n, &cpu);
  if (rc) return rc;
  if (cpu) { memcpy(cpu, in, n); __sync_synchronize(); return 0; }
  return cudaMemcpy((void*)gpu_va, in, n, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -EIO;
}
