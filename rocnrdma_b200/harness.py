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


def available_backend() -> str:
    """'dev' when the kernel module is loaded and its node is reachable, else 'user'."""
    return "dev" if os.path.exists(DEVICE_PATH) and os.access(DEVICE_PATH, os.R_OK | os.W_OK) else "user"
