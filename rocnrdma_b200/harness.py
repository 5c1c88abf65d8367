"""b200p2ptest user<->kernel ABI in Python (mirror of kmod/include/b200p2ptest.h).

Used three ways: against the real ``/dev/b200p2ptest`` (``DevBackend``), against the userspace
simulation of the kernel module (tests/test_kmod_sim.py), and by the CUDA-driver-API "twin"
(``UserBackend``) that offers the same four verbs without loading a module.  The reference ships
only the kernel half of this harness (include/amdp2ptest.h, tests/amdp2ptest.c); the program that
drives the ioctls was never published.
"""
from __future__ import annotations

import ctypes as C
import os

MAGIC = ord("B")
DEVICE_PATH = "/dev/b200p2ptest"
GPU_PAGE_SIZE = 1 << 16
MAX_BUS_ADDRS = 512

_IOC_NRBITS, _IOC_TYPEBITS, _IOC_SIZEBITS = 8, 8, 14
_IOC_NRSHIFT = 0
_IOC_TYPESHIFT = _IOC_NRSHIFT + _IOC_NRBITS
_IOC_SIZESHIFT = _IOC_TYPESHIFT + _IOC_TYPEBITS
_IOC_DIRSHIFT = _IOC_SIZESHIFT + _IOC_SIZEBITS
_IOC_WRITE, _IOC_READ = 1, 2


def _IOWR(type_, nr, struct) -> int:
    return ((_IOC_READ | _IOC_WRITE) << _IOC_DIRSHIFT) | (type_ << _IOC_TYPESHIFT) | (nr << _IOC_NRSHIFT) | \
        (C.sizeof(struct) << _IOC_SIZESHIFT)


class IsGpuAddress(C.Structure):
    _fields_ = [("addr", C.c_uint64), ("ret_value", C.c_uint32), ("reserved", C.c_uint32)]


class GetPageSize(C.Structure):
    _fields_ = [("addr", C.c_uint64), ("length", C.c_uint64), ("page_size", C.c_uint64)]


class GetPages(C.Structure):
    _fields_ = [("addr", C.c_uint64), ("length", C.c_uint64), ("handle", C.c_uint64), ("entries", C.c_uint32),
                ("page_size", C.c_uint32)]


class PutPages(C.Structure):
    _fields_ = [("addr", C.c_uint64), ("length", C.c_uint64), ("released", C.c_uint32), ("reserved", C.c_uint32)]


class GetBusAddrs(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("first", C.c_uint32), ("count", C.c_uint32), ("addrs", C.c_uint64 * MAX_BUS_ADDRS)]


IOCTL_IS_GPU_ADDRESS = _IOWR(MAGIC, 1, IsGpuAddress)
IOCTL_GET_PAGE_SIZE = _IOWR(MAGIC, 2, GetPageSize)
IOCTL_GET_PAGES = _IOWR(MAGIC, 3, GetPages)
IOCTL_PUT_PAGES = _IOWR(MAGIC, 4, PutPages)
IOCTL_GET_BUS_ADDRS = _IOWR(MAGIC, 5, GetBusAddrs)


class HarnessError(OSError):
    pass


class DevBackend:
    """Talks to the kernel module through /dev/b200p2ptest."""

    def __init__(self, path: str = DEVICE_PATH):
        self.fd = os.open(path, os.O_RDWR)

    def _ioctl(self, cmd, arg):
        import fcntl
        fcntl.ioctl(self.fd, cmd, arg)
        return arg

    def is_gpu_address(self, addr: int) -> bool:
        return bool(self._ioctl(IOCTL_IS_GPU_ADDRESS, IsGpuAddress(addr=addr)).ret_value)

    def get_page_size(self, addr: int, length: int) -> int:
        return self._ioctl(IOCTL_GET_PAGE_SIZE, GetPageSize(addr=addr, length=length)).page_size

    def get_pages(self, addr: int, length: int) -> GetPages:
        return self._ioctl(IOCTL_GET_PAGES, GetPages(addr=addr, length=length))

    def put_pages(self, addr: int, length: int) -> int:
        return self._ioctl(IOCTL_PUT_PAGES, PutPages(addr=addr, length=length)).released

    def bus_addrs(self, handle: int, first: int = 0, count: int = MAX_BUS_ADDRS):
        r = self._ioctl(IOCTL_GET_BUS_ADDRS, GetBusAddrs(handle=handle, first=first, count=count))
        return list(r.addrs[:r.count])

    def mmap(self, gpu_va: int, length: int):
        import mmap
        return mmap.mmap(self.fd, length, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE, offset=gpu_va)

    def close(self):
        if self.fd >= 0:
            os.close(self.fd)
            self.fd = -1


class UserBackend:
    """The same verbs through the CUDA driver API (csrc/reg/p2ptest_user.cc): no kernel module needed.
    A pin is a dma-buf export of the range.  The CPU window is an mmap() of that dma-buf where the exporting driver offers
    one (``map_window``: then peek / poke are CPU loads / stores through the BAR, as in the kernel harness); otherwise peek /
    poke fall back to cudaMemcpy, which proves nothing about the aperture -- ``window_kind`` says which it was."""

    def __init__(self):
        from . import _native as N
        self._lib = N.load()
        self._s = self._lib.rn_p2p_open()

    def is_gpu_address(self, addr: int) -> bool:
        return bool(self._lib.rn_p2p_is_gpu_address(self._s, addr))

    def get_page_size(self, addr: int, length: int) -> int:
        out = C.c_uint64()
        rc = self._lib.rn_p2p_get_page_size(self._s, addr, length, C.byref(out))
        if rc:
            raise HarnessError(-rc, os.strerror(-rc))
        return out.value

    def get_pages(self, addr: int, length: int) -> GetPages:
        h, e, p = C.c_uint64(), C.c_uint32(), C.c_uint32()
        rc = self._lib.rn_p2p_get_pages(self._s, addr, length, C.byref(h), C.byref(e), C.byref(p))
        if rc:
            raise HarnessError(-rc, os.strerror(-rc))
        return GetPages(addr=addr, length=length, handle=h.value, entries=e.value, page_size=p.value)

    def put_pages(self, addr: int, length: int) -> int:
        return self._lib.rn_p2p_put_pages(self._s, addr, length)

    def pin_size(self, handle: int) -> int:
        return self._lib.rn_p2p_pin_size(self._s, handle)

    @property
    def live_pins(self) -> int:
        return self._lib.rn_p2p_live_pins(self._s)

    def map_window(self, handle: int):
        """Try to map the pin into this process (mmap of its dma-buf).  Returns ``(cpu_address, length)`` or raises
        HarnessError with the errno of the attempt (driver 580: the NVIDIA exporter has no mmap op)."""
        a, n = C.c_uint64(), C.c_uint64()
        rc = self._lib.rn_p2p_mmap(self._s, handle, C.byref(a), C.byref(n))
        if rc:
            raise HarnessError(-rc, os.strerror(-rc))
        return a.value, n.value

    def window_kind(self, gpu_va: int) -> str:
        k = self._lib.rn_p2p_window_kind(self._s, gpu_va)
        return "dmabuf-mmap (CPU loads/stores through the BAR)" if k == 1 else "cudaMemcpy (no CPU mapping of the pin available)"

    def peek(self, gpu_va: int, n: int) -> bytes:
        buf = C.create_string_buffer(n)
        rc = self._lib.rn_p2p_peek(self._s, gpu_va, buf, n)
        if rc:
            raise HarnessError(-rc, os.strerror(-rc))
        return buf.raw

    def poke(self, gpu_va: int, data: bytes):
        rc = self._lib.rn_p2p_poke(self._s, gpu_va, data, len(data))
        if rc:
            raise HarnessError(-rc, os.strerror(-rc))

    def close(self) -> int:
        if self._s:
            n = self._lib.rn_p2p_close(self._s)
            self._s = None
            return n
        return 0


def open_backend(kind: str = "auto"):
    kind = available_backend() if kind == "auto" else kind
    return DevBackend() if kind == "dev" else UserBackend()


def available_backend() -> str:
    """'dev' when the kernel module is loaded and its node is reachable, else 'user'."""
    return "dev" if os.path.exists(DEVICE_PATH) and os.access(DEVICE_PATH, os.R_OK | os.W_OK) else "user"
