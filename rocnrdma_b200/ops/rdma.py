"""K1 / K2 / K6 wrappers: GPU-initiated RDMA streams and verification helpers."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List

import torch

from .. import _native as N

WAIT_STATUS = {0: "OK", -1: "TIMEOUT", -2: "CQE_ERROR", -3: "QP_ERROR"}


def _stream_ptr(stream=None) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def work_stream(ctx, stream=None):
    """Stream for engine-facing kernels: never the legacy default stream (see Context.stream)."""
    # native launchers use the calling thread's current device: make it this context's (creating a
    # second Context, or torch work on another GPU, may have changed it)
    N.load().rn_set_device(ctx.device)
    if getattr(ctx, "auto_sweep", True):
        ctx.sweep_revoked()               # driver-originated revocation check before a kernel may post (one driver query per MR)
    if stream is not None and int(stream.cuda_stream) != 0:
        return stream
    cur = torch.cuda.current_stream(ctx.device)
    if int(cur.cuda_stream) != 0:
        return cur
    ws = ctx.stream
    # The caller's tensors were produced on the legacy default stream (torch.zeros / randn / copies): order the work stream
    # after it, or the kernel races with their initialisation (compute-sanitizer's slower kernels made exactly that visible:
    # a zero-fill landing after the GEMM had written C).  Not while an engine is resident: anything recorded on the legacy
    # stream then queues behind the persistent kernel and the poster it gates would never start --
    # allocate and initialise before engine_start(), as the residency rules say.
    if not ctx.engine_running:
        ws.wait_stream(cur)
    return ws


@dataclass
class StreamResult:
    status: List[str]
    t_start_ns: List[int]
    t_end_ns: List[int]
    done: List[int]
    bytes_per_msg: int

    @property
    def ok(self) -> bool:
        return all(s == "OK" for s in self.status)

    @property
    def device_ns(self) -> int:
        """Device-timed span: first QP start to last QP end (%globaltimer)."""
        return max(self.t_end_ns) - min(self.t_start_ns)

    @property
    def gbps(self) -> float:
        return sum(self.done) * self.bytes_per_msg / max(self.device_ns, 1)

    @property
    def us_per_msg(self) -> float:
        return self.device_ns / 1e3 / max(max(self.done), 1)


def rdma_stream(qps, opcode: int, src_mr, dst_mr, nbytes: int, iters: int = 1, window: int = 0,
                signal_every: int = 1, burst: int = 1, stride: int = 0, slot_stride: int = 0, nslots: int = 1,
                timeout_ms: int = 2000, stream=None, sync: bool = True, out=None, post_only: bool = False):
    """Launch the device poster: one CTA per QP posts ``iters`` work requests of
    ``nbytes`` (window-limited), polls its CQ on the device and returns device times.
    ``burst`` work requests share one slot reservation and one doorbell (perftest ``--post_list``),
    ``signal_every`` is the CQ moderation (``--cq-mod``); both are clamped so the window can always drain.

    For RDMA_READ ``src_mr`` is the local destination and ``dst_mr`` the remote source,
    mirroring the laddr/raddr roles in the WQE.  ``post_only`` skips the final completion wait (``iters`` must fit
    the window): the work is reaped later -- by ``Context.engine_run_oneshot`` under a profiler, or by the caller.
    """
    lib = N.load()
    if not isinstance(qps, (list, tuple)):
        qps = [qps]
    ctx = qps[0].ctx
    ws = work_stream(ctx, stream)
    nq = len(qps)
    qp_arr = (C.c_uint64 * nq)(*[q.dev_ptr for q in qps])
    out_addr, out_view = ctx.scratch(nq * 64) if out is None else out
    rc = lib.rn_k_rdma_stream(_stream_ptr(ws), qp_arr, nq, opcode | (0x80000000 if post_only else 0), src_mr.addr, src_mr.lkey,
                              dst_mr.addr if dst_mr is not None else 0, dst_mr.rkey if dst_mr is not None else 0,
                              stride, nbytes, iters, window, signal_every, burst, slot_stride, nslots, timeout_ms, out_addr)
    if rc:
        raise N.NativeError(f"rdma_stream launch failed: cuda error {rc}")
    if not sync:
        return out_view, ws
    ws.synchronize()
    return parse_stream_out(out_view, nq, nbytes)


def parse_stream_out(out, nqp: int, nbytes: int) -> StreamResult:
    words = (C.c_int64 * (nqp * 8)).from_buffer(out)
    o = [list(words[i * 8:(i + 1) * 8]) for i in range(nqp)]
    return StreamResult(status=[WAIT_STATUS.get(r[0], str(r[0])) for r in o], t_start_ns=[r[1] for r in o],
                        t_end_ns=[r[2] for r in o], done=[r[3] for r in o], bytes_per_msg=nbytes)


def fill_random(t: torch.Tensor, seed: int = 1, stream=None):
    N.load().rn_k_fill_random(_stream_ptr(stream), t.data_ptr(), t.numel() * t.element_size(), seed)
    return t


def fill_bf16(t: torch.Tensor, seed: int = 1, scale: float = 1.0, stream=None):
    assert t.dtype == torch.bfloat16
    N.load().rn_k_fill_bf16(_stream_ptr(stream), t.data_ptr(), t.numel(), seed, scale)
    return t


def checksum(t: torch.Tensor, stream=None) -> int:
    out = torch.zeros(1, dtype=torch.int64, device=t.device)
    N.load().rn_k_checksum(_stream_ptr(stream), t.data_ptr(), t.numel() * t.element_size(), out.data_ptr())
    return int(out.item()) & 0xFFFFFFFFFFFFFFFF


def compare(a: torch.Tensor, b: torch.Tensor, nbytes=None, stream=None) -> int:
    """Number of mismatching 16-byte words (aligned) or bytes (unaligned)."""
    n = a.numel() * a.element_size() if nbytes is None else nbytes
    out = torch.zeros(1, dtype=torch.int64, device=a.device)
    N.load().rn_k_compare(_stream_ptr(stream), a.data_ptr(), b.data_ptr(), n, out.data_ptr())
    return int(out.item())


def l2_flush(scratch: torch.Tensor, value: int = 0, stream=None):
    N.load().rn_k_l2_flush(_stream_ptr(stream), scratch.data_ptr(), scratch.numel() * scratch.element_size(), value)


def recv_consume(qp, n: int, max_imm: int = 0, stamps=None, timeout_ms: int = 2000, stream=None, sync: bool = True,
                 scratch_off: int = 8192, prepost_mr=None, prepost_bytes: int = 0):
    """Receive-side consumer kernel on ``qp``'s GPU: waits for ``n`` arrivals (SEND / RDMA_WRITE_IMM) on the
    receive CQ and, if ``stamps`` (int64 tensor on that GPU) is given, stamps %globaltimer per immediate."""
    ctx = qp.ctx
    ws = work_stream(ctx, stream)
    out_addr, out_view = ctx.scratch(64, offset=scratch_off)
    # prepost_mr: the kernel posts its own n receive WQEs from the device before polling (buffer i at offset i * prepost_bytes)
    rc = N.load().rn_k_recv_consume(_stream_ptr(ws), qp.dev_ptr, n, max_imm, stamps.data_ptr() if stamps is not None else 0,
                                    out_addr, timeout_ms, prepost_mr.addr if prepost_mr is not None else 0,
                                    prepost_mr.lkey if prepost_mr is not None else 0,
                                    prepost_bytes if prepost_mr is not None else 0xFFFFFFFF)
    if rc:
        raise N.NativeError(f"recv_consume launch failed ({rc})")
    if not sync:
        return out_view, ws
    ws.synchronize()
    return parse_recv(out_view)


def parse_recv(view) -> dict:
    w = (C.c_int64 * 8).from_buffer(view)
    return dict(status=WAIT_STATUS.get(w[0], str(w[0])), device_ns=w[2] - w[1], t_start_ns=w[1], t_end_ns=w[2], seen=w[3], bytes=w[4])


def shared_post_stress(qp, src_mr, dst_mr, ctas: int = 64, per_cta: int = 64, timeout_ms: int = 3000, stream=None):
    """``ctas`` CTAs hammer one QP with ``per_cta`` 64-byte writes each through the shared submit."""
    ctx = qp.ctx
    ws = work_stream(ctx, stream)
    out_addr, out_view = ctx.scratch(64, offset=12288)
    counter = ctx.dev_scratch(64, offset=900 << 10)
    rc = N.load().rn_k_shared_post_stress(_stream_ptr(ws), qp.dev_ptr, ctas, src_mr.addr, src_mr.lkey, dst_mr.addr, dst_mr.rkey,
                                          per_cta, counter, out_addr, timeout_ms)
    if rc:
        raise N.NativeError(f"shared_post_stress launch failed ({rc})")
    ws.synchronize()
    w = (C.c_int64 * 8).from_buffer(out_view)
    return dict(status=WAIT_STATUS.get(w[0], str(w[0])), device_ns=w[2] - w[1], posted=w[3])
