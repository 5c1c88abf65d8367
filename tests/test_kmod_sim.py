"""Kernel modules under the userspace simulation (kmod/shim): the 10 behaviours of SURVEY.md section 4.2
for both the PeerDirect bridge (b200p2p.c) and the harness (b200p2ptest.c), plus the orderings the
reference's bare-flag revoke handling could not distinguish."""
import ctypes as C
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rocnrdma_b200 import harness as H  # noqa: E402
from tools import build_kmod_sim  # noqa: E402

VA = 0x7F00_0000_0000
PAGE = 1 << 16
u64 = C.c_uint64


@pytest.fixture(scope="module")
def lib():
    l = C.CDLL(str(build_kmod_sim.build()))
    for name, res, args in [
        ("sim_gpu_alloc", C.c_int, [u64, u64]), ("sim_gpu_free", C.c_int, [u64]), ("sim_gpu_bus_addr", u64, [u64]),
        ("sim_ib_reg_mr", C.c_long, [u64, u64, C.c_int]), ("sim_ib_dereg_mr", C.c_int, [C.c_long]),
        ("sim_ib_mr_dma", C.c_int, [C.c_long, C.c_int, C.POINTER(u64), C.POINTER(u64)]),
        ("sim_ib_mr_page_size", u64, [C.c_long]), ("sim_ib_bad_sequence", C.c_int, [C.c_int, u64, u64]),
        ("sim_dev_open", C.c_void_p, []), ("sim_dev_ioctl", C.c_long, [C.c_void_p, C.c_uint, C.c_void_p]),
        ("sim_dev_close", C.c_int, [C.c_void_p]), ("sim_dev_mmap", C.c_int, [C.c_void_p, u64, u64, C.POINTER(u64), C.c_int]),
        ("sim_live_allocs", C.c_long, []), ("sim_ib_client_name", C.c_char_p, []), ("sim_ib_client_version", C.c_char_p, []),
        ("sim_dev_name", C.c_char_p, []), ("sim_debugfs_read", C.c_int, [C.c_char_p, C.c_char_p, C.c_int]),
        ("sim_param_b200p2p_debug", None, [C.c_long]), ("sim_param_b200p2p_max_pin_mb", None, [C.c_long]),
        ("sim_param_b200p2p_enable", None, [C.c_long]),
        ("sim_nv_revoke_during_dma_map", None, [u64]), ("sim_ib_set_release_in_invalidate", None, [C.c_int]),
    ]:
        getattr(l, name).restype, getattr(l, name).argtypes = res, args
    return l


@pytest.fixture
def bridge(lib):
    lib.sim_reset()
    base = lib.sim_live_allocs()
    assert lib.sim_b200p2p_load() == 0
    yield lib
    lib.sim_b200p2p_unload()
    assert lib.sim_live_page_tables() == 0 and lib.sim_live_dma_mappings() == 0, "pin or mapping leaked"
    assert lib.sim_nv_misuse() == 0, "NVIDIA P2P interface misused (put after revoke / double free)"
    assert lib.sim_lock_errors() == 0
    assert lib.sim_module_refcount() == 0
    assert lib.sim_live_allocs() == base, "kernel allocation leaked"


@pytest.fixture
def dev(lib):
    lib.sim_reset()
    base = lib.sim_live_allocs()
    assert lib.sim_b200p2ptest_load() == 0
    yield lib
    lib.sim_b200p2ptest_unload()
    assert lib.sim_live_page_tables() == 0 and lib.sim_nv_misuse() == 0 and lib.sim_lock_errors() == 0
    assert lib.sim_live_allocs() == base


def ioctl(lib, f, cmd, arg):
    return lib.sim_dev_ioctl(f, cmd, C.byref(arg))


# =============================================================== bridge (amdp2p.c counterpart)
def test_registers_under_its_name(bridge):
    assert bridge.sim_ib_client_name() == b"b200p2p" and bridge.sim_ib_client_version() == b"1.0"


def test_load_fails_when_ib_core_refuses(lib):               # behaviour 10 (amdp2p.c:393-396)
    lib.sim_reset()
    lib.sim_ib_set_refuse(1)
    assert lib.sim_b200p2p_load() == -22
    assert lib.sim_log_count(0) >= 1


def test_host_address_is_not_claimed(bridge):                # behaviour 1
    bridge.sim_gpu_alloc(VA, 4 * PAGE)
    assert bridge.sim_ib_reg_mr(0x5555_0000_0000, PAGE, 0) == -95      # acquire answered "not mine"
    assert bridge.sim_module_refcount() == 0


def test_reg_mr_pins_and_hands_out_bus_addresses(bridge):    # behaviours 1-3
    bridge.sim_gpu_alloc(VA, 8 * PAGE)
    mr = bridge.sim_ib_reg_mr(VA + PAGE, 3 * PAGE, 2)
    assert mr >= 0
    assert bridge.sim_ib_mr_nmap(mr) == 3 and bridge.sim_ib_mr_page_size(mr) == PAGE
    a, n = u64(), u64()
    for i in range(3):
        assert bridge.sim_ib_mr_dma(mr, i, C.byref(a), C.byref(n)) == 0
        # per-HCA mapping: device id 2 is folded into the IOVA by the mock (the reference ignores dma_device)
        assert a.value == bridge.sim_gpu_bus_addr(VA + (1 + i) * PAGE) + (2 << 52) and n.value == PAGE
    assert bridge.sim_module_refcount() == 1                 # behaviour 9: no rmmod while registered
    assert bridge.sim_ib_dereg_mr(mr) == 0
    assert bridge.sim_live_pins() == 0


def test_unaligned_range_is_widened_to_gpu_pages(bridge):
    bridge.sim_gpu_alloc(VA, 8 * PAGE)
    mr = bridge.sim_ib_reg_mr(VA + 100, PAGE, 0)             # straddles two 64 KiB pages
    assert mr >= 0 and bridge.sim_ib_mr_nmap(mr) == 2
    bridge.sim_ib_dereg_mr(mr)


@pytest.mark.parametrize("which,errno", [(0, -22), (1, -22), (2, -22), (3, -22), (5, -22)])
def test_malformed_callback_sequences(bridge, which, errno):  # behaviours 7 and 8
    bridge.sim_gpu_alloc(VA, 8 * PAGE)
    assert bridge.sim_ib_bad_sequence(which, VA, 2 * PAGE) == errno


def test_release_with_live_pin_cleans_up(bridge):
    bridge.sim_gpu_alloc(VA, 8 * PAGE)
    assert bridge.sim_ib_bad_sequence(4, VA, 2 * PAGE) == 0
    assert bridge.sim_log_count(1) >= 1                      # warned about it


@pytest.mark.parametrize("sync_invalidate", [1, 0])
def test_revocation_invalidates_mr_and_later_put_is_a_noop(bridge, sync_invalidate):   # behaviour 6
    bridge.sim_ib_set_sync_invalidate(sync_invalidate)
    bridge.sim_gpu_alloc(VA, 4 * PAGE)
    mr = bridge.sim_ib_reg_mr(VA, 4 * PAGE, 0)
    assert mr >= 0
    assert bridge.sim_gpu_free(VA) == 1                      # cudaFree under a live MR -> one callback
    assert bridge.sim_ib_invalidate_calls() == 1 and bridge.sim_ib_mr_invalidated(mr) == 1
    assert bridge.sim_live_page_tables() == 0 and bridge.sim_live_dma_mappings() == 0    # freed by the revoke path
    assert bridge.sim_ib_dereg_mr(mr) == 0                   # later dma_unmap/put_pages must not touch the dead pin
    assert bridge.sim_nv_misuse() == 0


def test_revocation_of_one_mr_leaves_others_alone(bridge):
    bridge.sim_gpu_alloc(VA, 4 * PAGE)
    bridge.sim_gpu_alloc(VA + 16 * PAGE, 4 * PAGE)
    a = bridge.sim_ib_reg_mr(VA, 4 * PAGE, 0)
    b = bridge.sim_ib_reg_mr(VA + 16 * PAGE, 4 * PAGE, 1)
    bridge.sim_gpu_free(VA)
    assert bridge.sim_ib_mr_invalidated(a) == 1 and bridge.sim_ib_mr_invalidated(b) == 0
    addr, n = u64(), u64()
    assert bridge.sim_ib_mr_dma(b, 0, C.byref(addr), C.byref(n)) == 0
    bridge.sim_ib_dereg_mr(a)
    bridge.sim_ib_dereg_mr(b)


def test_memory_freed_during_registration(bridge):
    bridge.sim_gpu_alloc(VA, 4 * PAGE)
    bridge.sim_nv_fail_next_get_pages(0)
    # the first get_pages is acquire's probe pin; arm the early revoke for the real one
    mr = bridge.sim_ib_reg_mr(VA, PAGE, 0)
    assert mr >= 0
    bridge.sim_ib_dereg_mr(mr)
    bridge.sim_nv_revoke_during_get_pages(1)
    rc = bridge.sim_ib_reg_mr(VA, PAGE, 0)
    assert rc in (-14, -95)          # either the probe or the real pin saw the memory disappear; nothing leaks (fixture)


def test_driver_failures_propagate(bridge):
    bridge.sim_gpu_alloc(VA, 4 * PAGE)
    bridge.sim_nv_fail_next_get_pages(2)                     # probe ok, real pin fails
    assert bridge.sim_ib_reg_mr(VA, PAGE, 0) == -12
    bridge.sim_nv_fail_next_dma_map(1)
    assert bridge.sim_ib_reg_mr(VA, PAGE, 0) == -5
    assert bridge.sim_live_pins() == 0


# =============================================================== harness (amdp2ptest.c counterpart)
def test_ioctl_numbers_encode_the_struct_size():
    # the reference encodes sizeof(pointer) (include/amdp2ptest.h:62-72); ours carry the real struct
    assert (H.IOCTL_GET_PAGES >> 16) & 0x3FFF == C.sizeof(H.GetPages) == 32
    assert (H.IOCTL_IS_GPU_ADDRESS >> 16) & 0x3FFF == 16 and H.IOCTL_IS_GPU_ADDRESS >> 30 == 3   # _IOWR
    assert len({H.IOCTL_IS_GPU_ADDRESS, H.IOCTL_GET_PAGE_SIZE, H.IOCTL_GET_PAGES, H.IOCTL_PUT_PAGES, H.IOCTL_GET_BUS_ADDRS}) == 5


def test_device_identity_and_mode(dev):
    assert dev.sim_dev_name() == b"b200p2ptest"
    assert dev.sim_dev_mode() == 0o660                       # not 0777 (tests/amdp2ptest.c:427)


def test_unknown_ioctl_is_einval(dev):
    f = dev.sim_dev_open()
    assert dev.sim_dev_ioctl(f, 0xDEAD, None) == -22
    dev.sim_dev_close(f)


def test_is_gpu_address_and_page_size(dev):                  # behaviours 1 and 2
    dev.sim_gpu_alloc(VA, 4 * PAGE)
    f = dev.sim_dev_open()
    q = H.IsGpuAddress(addr=VA + 12345)
    assert ioctl(dev, f, H.IOCTL_IS_GPU_ADDRESS, q) == 0 and q.ret_value == 1
    q = H.IsGpuAddress(addr=0x1000)
    assert ioctl(dev, f, H.IOCTL_IS_GPU_ADDRESS, q) == 0 and q.ret_value == 0
    s = H.GetPageSize(addr=VA, length=2 * PAGE)
    assert ioctl(dev, f, H.IOCTL_GET_PAGE_SIZE, s) == 0 and s.page_size == PAGE
    s = H.GetPageSize(addr=0x1000, length=4096)
    assert ioctl(dev, f, H.IOCTL_GET_PAGE_SIZE, s) == -14    # -EFAULT like the reference (:189-192)
    dev.sim_dev_close(f)
    assert dev.sim_live_pins() == 0


def test_pin_bus_addresses_and_full_mmap(dev):               # behaviour 3
    dev.sim_gpu_alloc(VA, 8 * PAGE)
    f = dev.sim_dev_open()
    g = H.GetPages(addr=VA + PAGE, length=4 * PAGE)
    assert ioctl(dev, f, H.IOCTL_GET_PAGES, g) == 0 and g.entries == 4 and g.page_size == PAGE and g.handle
    b = H.GetBusAddrs(handle=g.handle, first=1, count=8)
    assert ioctl(dev, f, H.IOCTL_GET_BUS_ADDRS, b) == 0 and b.count == 3
    assert list(b.addrs[:3]) == [dev.sim_gpu_bus_addr(VA + (2 + i) * PAGE) for i in range(3)]
    # mmap a window that starts mid-page and spans three GPU pages: every page at its own bus address
    out = (u64 * 30)()
    n = dev.sim_dev_mmap(f, VA + PAGE + 4096, 2 * PAGE, out, 10)
    assert n == 3
    triples = [(out[3 * i], out[3 * i + 1], out[3 * i + 2]) for i in range(n)]
    assert triples[0] == (0, dev.sim_gpu_bus_addr(VA + PAGE) + 4096, PAGE - 4096)
    assert triples[1] == (PAGE - 4096, dev.sim_gpu_bus_addr(VA + 2 * PAGE), PAGE)
    assert triples[2] == (2 * PAGE - 4096, dev.sim_gpu_bus_addr(VA + 3 * PAGE), 4096)
    assert dev.sim_dev_mmap(f, VA, 2 * PAGE, out, 10) == -22             # not inside the pin
    assert dev.sim_dev_mmap(f, VA + 4 * PAGE, 2 * PAGE, out, 10) == -22  # runs past its end
    dev.sim_dev_close(f)


def test_misaligned_pin_is_rejected(dev):
    dev.sim_gpu_alloc(VA, 4 * PAGE)
    f = dev.sim_dev_open()
    assert ioctl(dev, f, H.IOCTL_GET_PAGES, H.GetPages(addr=VA + 4096, length=PAGE)) == -22
    assert ioctl(dev, f, H.IOCTL_GET_PAGES, H.GetPages(addr=VA, length=PAGE + 4096)) == -22
    assert ioctl(dev, f, H.IOCTL_GET_PAGES, H.GetPages(addr=0x10000, length=PAGE)) == -14   # host memory
    dev.sim_dev_close(f)


def test_double_pin_one_put_releases_all(dev):               # behaviour 4 (tests/amdp2ptest.c:296-299)
    dev.sim_gpu_alloc(VA, 4 * PAGE)
    f = dev.sim_dev_open()
    for _ in range(3):
        assert ioctl(dev, f, H.IOCTL_GET_PAGES, H.GetPages(addr=VA, length=2 * PAGE)) == 0
    assert ioctl(dev, f, H.IOCTL_GET_PAGES, H.GetPages(addr=VA, length=PAGE)) == 0       # different length: survives
    assert dev.sim_live_pins() == 4
    p = H.PutPages(addr=VA, length=2 * PAGE)
    assert ioctl(dev, f, H.IOCTL_PUT_PAGES, p) == 0 and p.released == 3
    assert dev.sim_live_pins() == 1
    p = H.PutPages(addr=VA + PAGE, length=PAGE)
    assert ioctl(dev, f, H.IOCTL_PUT_PAGES, p) == 0 and p.released == 0                  # no match is not an error (:303)
    dev.sim_dev_close(f)


def test_close_with_leaks_unpins_everything(dev):            # behaviour 5
    dev.sim_gpu_alloc(VA, 8 * PAGE)
    f1, f2 = dev.sim_dev_open(), dev.sim_dev_open()
    ioctl(dev, f1, H.IOCTL_GET_PAGES, H.GetPages(addr=VA, length=PAGE))
    ioctl(dev, f1, H.IOCTL_GET_PAGES, H.GetPages(addr=VA + PAGE, length=2 * PAGE))
    ioctl(dev, f2, H.IOCTL_GET_PAGES, H.GetPages(addr=VA, length=PAGE))
    assert dev.sim_live_pins() == 3
    dev.sim_dev_close(f1)
    assert dev.sim_live_pins() == 1                           # per-fd lists: f2's pin is untouched
    dev.sim_dev_close(f2)
    assert dev.sim_live_pins() == 0


def test_revocation_is_loud_and_node_becomes_stale(dev):     # behaviour 6
    dev.sim_gpu_alloc(VA, 4 * PAGE)
    f = dev.sim_dev_open()
    g = H.GetPages(addr=VA, length=2 * PAGE)
    ioctl(dev, f, H.IOCTL_GET_PAGES, g)
    errs = dev.sim_log_count(0)
    assert dev.sim_gpu_free(VA) == 1
    assert dev.sim_log_count(0) == errs + 1                   # logged at ERR level, like the reference (:81-82)
    assert ioctl(dev, f, H.IOCTL_GET_BUS_ADDRS, H.GetBusAddrs(handle=g.handle, count=4)) == -116   # -ESTALE
    out = (u64 * 30)()
    assert dev.sim_dev_mmap(f, VA, PAGE, out, 10) == -22
    p = H.PutPages(addr=VA, length=2 * PAGE)
    assert ioctl(dev, f, H.IOCTL_PUT_PAGES, p) == 0 and p.released == 1    # frees the node, no put_pages on a dead pin
    dev.sim_dev_close(f)


@pytest.mark.parametrize("nth", [0, 1])
def test_copy_faults_leak_nothing(dev, nth):                 # the reference leaks its node here (:243-253)
    dev.sim_gpu_alloc(VA, 4 * PAGE)
    f = dev.sim_dev_open()
    dev.sim_set_copy_fault(nth)                               # 0: copy_from_user fails, 1: copy_to_user fails
    assert ioctl(dev, f, H.IOCTL_GET_PAGES, H.GetPages(addr=VA, length=PAGE)) == -14
    assert dev.sim_live_pins() == 0
    dev.sim_dev_close(f)


def test_misc_register_failure_propagates(lib):              # behaviour 10
    lib.sim_reset()
    lib.sim_misc_set_fail(1)
    assert lib.sim_b200p2ptest_load() == -16


# =============================================================== round-2 hardening
def _stats(lib):
    buf = C.create_string_buffer(512)
    n = lib.sim_debugfs_read(b"stats", buf, 512)
    assert n > 0
    return {k: int(v) for k, v in (line.split() for line in buf.value.decode().strip().splitlines())}


def test_counters_are_browsable_in_debugfs_while_loaded(bridge):
    bridge.sim_gpu_alloc(VA, 8 * PAGE)
    a = bridge.sim_ib_reg_mr(VA, 2 * PAGE, 0)
    b = bridge.sim_ib_reg_mr(VA + 4 * PAGE, PAGE, 1)
    st = _stats(bridge)
    assert st["acquired"] == 2 and st["pinned"] == 2 and st["mapped"] == 2 and st["live"] == 2 and st["revoked"] == 0
    bridge.sim_gpu_free(VA)
    bridge.sim_ib_dereg_mr(a)
    bridge.sim_ib_dereg_mr(b)
    st = _stats(bridge)
    assert st["revoked"] == 2 and st["released"] == 2 and st["live"] == 0


def test_enable_parameter_stops_new_claims_and_leaves_live_registrations_alone(bridge):
    """enable=0: the client stays registered but claims nothing new; what is registered keeps translating, is revoked by the
    driver and is released as usual.  Back to 1: claims resume."""
    bridge.sim_gpu_alloc(VA, 8 * PAGE)
    live = bridge.sim_ib_reg_mr(VA, 2 * PAGE, 0)
    assert live >= 0
    bridge.sim_param_b200p2p_enable(0)
    try:
        assert bridge.sim_ib_reg_mr(VA + 4 * PAGE, PAGE, 0) == -95       # nobody claims the range: ibv_reg_mr fails cleanly
        assert bridge.sim_live_pins() == 1                              # the probe pin never happened
        bridge.sim_gpu_free(VA)                                         # the live registration is still revoked by the driver ...
        assert _stats(bridge)["revoked"] == 1
        bridge.sim_ib_dereg_mr(live)                                    # ... and released normally
        assert _stats(bridge)["live"] == 0 and bridge.sim_live_pins() == 0
    finally:
        bridge.sim_param_b200p2p_enable(1)
    bridge.sim_gpu_alloc(VA, 8 * PAGE)
    mr = bridge.sim_ib_reg_mr(VA, PAGE, 0)
    assert mr >= 0
    bridge.sim_ib_dereg_mr(mr)


def test_module_parameters(bridge):
    bridge.sim_gpu_alloc(VA, 64 * PAGE)
    bridge.sim_param_b200p2p_max_pin_mb(1)                      # 1 MiB = 16 GPU pages
    assert bridge.sim_ib_reg_mr(VA, 32 * PAGE, 0) == -95         # not claimed: ibv_reg_mr fails cleanly, nothing pinned
    assert bridge.sim_live_pins() == 0 and _stats(bridge)["refused"] == 1
    mr = bridge.sim_ib_reg_mr(VA, 16 * PAGE, 0)
    assert mr >= 0
    bridge.sim_param_b200p2p_max_pin_mb(0)
    bridge.sim_ib_dereg_mr(mr)
    # debug=1 turns the pr_debug breadcrumbs into INFO lines (kernels without dynamic debug)
    info = bridge.sim_log_count(2)
    bridge.sim_param_b200p2p_debug(1)
    mr = bridge.sim_ib_reg_mr(VA, PAGE, 0)
    bridge.sim_ib_dereg_mr(mr)
    bridge.sim_param_b200p2p_debug(0)
    assert bridge.sim_log_count(2) > info


def test_release_from_inside_the_invalidate_upcall(bridge):
    """ib_core destroys the MR -- release included -- while the module's free callback is still on the stack: the
    callback returns into a context that has already been released (kref keeps it alive until then)."""
    bridge.sim_ib_set_release_in_invalidate(1)
    bridge.sim_gpu_alloc(VA, 4 * PAGE)
    mr = bridge.sim_ib_reg_mr(VA, 4 * PAGE, 0)
    assert mr >= 0
    assert bridge.sim_gpu_free(VA) == 1
    assert _stats(bridge)["live"] == 0 and _stats(bridge)["released"] == 1
    assert bridge.sim_module_refcount() == 0
    assert bridge.sim_ib_dereg_mr(mr) == 0                       # nothing left to do, and nothing is done twice


@pytest.mark.parametrize("release_inside", [0, 1])
def test_revoke_while_dma_map_is_inside_the_driver(bridge, release_inside):
    """The window the advisor flagged: dma_map has dropped its lock and handed the page table to
    nvidia_p2p_dma_map_pages() when the memory is freed.  The page table must stay alive until the call returns."""
    bridge.sim_ib_set_release_in_invalidate(release_inside)
    bridge.sim_gpu_alloc(VA, 4 * PAGE)
    bridge.sim_nv_revoke_during_dma_map(VA)
    rc = bridge.sim_ib_reg_mr(VA, 2 * PAGE, 0)
    assert rc in (-22, -14)                                      # the registration fails; the fixture checks nothing leaked
    assert bridge.sim_nv_misuse() == 0 and _stats(bridge)["live"] == 0


def test_harness_logs_every_ioctl_and_its_symbols_at_load(dev):
    """Parity with the reference's only observability (tests/amdp2ptest.c:145-350, 441-445)."""
    assert dev.sim_log_count(2) >= 2                             # the symbol dump and the "ready" line
    dev.sim_gpu_alloc(VA, 4 * PAGE)
    f = dev.sim_dev_open()
    before = dev.sim_log_count(2)
    ioctl(dev, f, H.IOCTL_IS_GPU_ADDRESS, H.IsGpuAddress(addr=VA))
    ioctl(dev, f, H.IOCTL_GET_PAGE_SIZE, H.GetPageSize(addr=VA, length=PAGE))
    ioctl(dev, f, H.IOCTL_GET_PAGES, H.GetPages(addr=VA, length=PAGE))
    ioctl(dev, f, H.IOCTL_PUT_PAGES, H.PutPages(addr=VA, length=PAGE))
    assert dev.sim_log_count(2) - before == 4
    dev.sim_dev_close(f)
    assert dev.sim_log_count(2) - before == 5


@pytest.mark.parametrize("sanitizer", ["address", "thread"])
def test_multithreaded_lifecycle_stress(sanitizer):
    """Registrars || GPU-driver revocations || harness users on real threads, three ib_core teardown orders, as one
    sanitizer-instrumented executable (kmod/tests/sim_stress.c)."""
    import os
    import subprocess
    exe = build_kmod_sim.build_stress(sanitizer)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", TSAN_OPTIONS="halt_on_error=1")
    env.pop("LD_PRELOAD", None)          # `make check-sanitize` preloads libasan for the shared-library runs; these are whole programs
    r = subprocess.run([str(exe), "0.7"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "STRESS OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ThreadSanitizer" not in r.stderr and "AddressSanitizer" not in r.stderr
