#!/usr/bin/env python
"""Headline benchmark: GPU-initiated RDMA write GB/s on HBM buffers, each GPU driving its own HCA.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (torchrun for N > 1) prints ONE
JSON line from rank 0.  One *step* = one GPU-initiated RDMA WRITE of ``--msg-bytes`` (default 256 MiB,
larger than the 126 MB L2, rotating over 4 distinct buffers) on every rank: an sm_100a kernel builds
the mlx5 WQE, rings the doorbell, and polls the CQ on the device; the software HCA's persistent DMA
engine (TMA bulk copies) moves the bytes HBM -> HBM.  ``value`` is the whole-job aggregate GB/s,
device-timed with CUDA events, max over ranks.

The box exposes no /dev/infiniband to the container (gpurun probe: HCAs visible in sysfs only, no
rdma-core), so the wire is the software HCA (``config.wire``), not a ConnectX-7; BASELINE.md publishes
no reference number, hence ``vs_baseline: null``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def reference_arm():
    # `pip install --no-index ... /root/reference` fails: the reference is a Linux kernel module for
    # AMD KFD + MLNX_OFED 3.2 (no setup.py / pyproject, no userspace, needs amd_rdma.h); see DESIGN.md.
    if int(os.environ.get("RANK", "0")) != 0:      # launched like our own arm (torchrun for N > 1): one line, from rank 0
        return 0
    print(json.dumps({"impl": "reference",
                      "unavailable": "reference is an AMD-KFD/MLNX_OFED kernel module (amdp2p.ko): not pip-installable, "
                                     "no userspace entry point, cannot build or load on a B200 box"}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--msg-bytes", type=int, default=256 << 20)
    ap.add_argument("--engine-ctas", type=int, default=128)
    ap.add_argument("--pipeline", type=int, default=8, help="steps in flight: step i is posted on QP/stream i %% pipeline")
    ap.add_argument("--extras", type=int, default=1, help="also run the fused-pack and small-message extras (untimed region)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm()

    import torch
    import rocnrdma_b200 as rn
    from rocnrdma_b200 import ops, wire as W
    from rocnrdma_b200.ops import pack as P
    from rocnrdma_b200.utils.clocks import ClockSampler
    from rocnrdma_b200.utils import roofline as R

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        # the GPU-initiated path has no CPU fallback by design: say so in one line instead of a traceback
        if rank == 0:
            print(json.dumps({"metric": "rdma_write_gbps_gpu_hbm_device_timed", "value": None, "impl": "ours",
                              "unavailable": "no CUDA device visible: the data path is sm_100a kernels only"}))
        return 2
    torch.cuda.set_device(local_rank)
    from rocnrdma_b200.utils.affinity import bind_to_gpu
    cpus = bind_to_gpu(local_rank)          # pinned buffers get first-touched on the GPU's NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local_rank])

    msg = args.msg_bytes
    nslots = 4
    dev = torch.device("cuda", local_rank)
    ctx = rn.Context(device=local_rank)
    src = torch.empty(msg * nslots, dtype=torch.uint8, device=dev)
    dst = torch.empty(msg * nslots, dtype=torch.uint8, device=dev)
    ops.fill_random(src, seed=1000 + rank)
    host_in = torch.empty(msg, dtype=torch.uint8).pin_memory()       # e2e: step inputs live in pinned host memory
    host_in.random_(0, 255)
    ms, md = ctx.reg_mr(src), ctx.reg_mr(dst)
    depth = max(1, min(args.pipeline, 8))
    qps = [ctx.loopback_qp(depth=64) for _ in range(depth)]   # step i goes to QP / stream i % depth: a few steps in flight,
    qp = qps[0]                                               # like the tx-depth of ib_write_bw (one WQE and one launch per step)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev_tail = [torch.cuda.Event() for _ in range(depth)]
    torch.cuda.synchronize()
    barrier()

    strs = ctx.streams(depth)
    stream = strs[0]
    outs = [ctx.scratch(64, 64 * i) for i in range(depth)]
    out_a = outs[0]
    mr_slot = [rn.api.MemoryRegion(ctx, ms.addr + i * msg, msg, ms.key, ms.access) for i in range(nslots)]
    md_slot = [rn.api.MemoryRegion(ctx, md.addr + i * msg, msg, md.key, md.access) for i in range(nslots)]

    def step(i, sync=False, pingpong=True):
        s = i % nslots
        j = (i % depth) if pingpong else 0
        return ops.rdma_stream(qps[j], W.OP_RDMA_WRITE, mr_slot[s], md_slot[s], msg, iters=1, stream=strs[j], sync=sync, out=outs[j])

    # ---------------- device-timed headline
    barrier()
    torch.cuda.synchronize()
    ctx.engine_start(ctas=args.engine_ctas, idle_timeout_ms=8000)
    for i in range(warmup):
        r = step(i, sync=True)
        assert r.ok, r.status
    sampler = ClockSampler(gpu_index=None, period_s=0.05).start() if rank == 0 else None
    for st in strs:
        st.synchronize()
    ev[0].record(stream)
    for st in strs[1:]:
        st.wait_event(ev[0])
    for i in range(args.steps):
        step(i)
    for j in range(1, depth):
        ev_tail[j].record(strs[j])
        stream.wait_event(ev_tail[j])
    ev[1].record(stream)
    ev[1].synchronize()
    dev_ms = ev[0].elapsed_time(ev[1])
    for o in outs[:min(depth, args.steps)]:
        last = ops.rdma.parse_stream_out(o[1], 1, msg)
        assert last.ok, last.status
    # H2D alone, for the e2e breakdown
    with torch.cuda.stream(stream):
        src[:msg].copy_(host_in, non_blocking=True)
        ev[4].record()
        for i in range(4):
            src[(i % nslots) * msg:(i % nslots + 1) * msg].copy_(host_in, non_blocking=True)
        ev[5].record()
    ev[5].synchronize()
    h2d_gbps = 4 * msg / (ev[4].elapsed_time(ev[5]) * 1e-3) / 1e9

    # ---------------- end to end through the public API: pinned host -> H2D -> GPU-posted write -> status D2H
    with torch.cuda.stream(stream):
        for i in range(2):
            src[:msg].copy_(host_in, non_blocking=True)
            step(0, pingpong=False)
        ev[2].record()
        for i in range(args.steps):
            s = i % nslots
            src[s * msg:(s + 1) * msg].copy_(host_in, non_blocking=True)     # H2D of this step's input
            step(i, pingpong=False)
            stream.synchronize()                                             # result visible to the host:
            st = ops.rdma.parse_stream_out(out_a[1], 1, msg)                 # 64 B status/timing words the kernel
            assert st.ok, st.status                                          # wrote to mapped pinned memory
        ev[3].record()
    ev[3].synchronize()
    e2e_ms = ev[2].elapsed_time(ev[3])
    clocks = sampler.stop() if sampler else {}

    extras = {}
    if args.extras:
        # fused bf16 -> fp8 pack + GPU-initiated write (config 3), and a small-message point
        n_el = min(1 << 28, (msg * nslots) // 2)
        n_el -= n_el % (1 << 22)
        if n_el >= (1 << 22):
            x = src[:2 * n_el].view(torch.bfloat16)
            chunk = 1 << 22
            nb = P.staging_bytes(n_el, chunk)
            stg = rn.api.MemoryRegion(ctx, md.addr, nb, md.key, md.access)
            rmt = rn.api.MemoryRegion(ctx, md.addr + (msg * nslots) // 2, nb, md.key, md.access)
            if 2 * nb <= msg * nslots:
                ctx.engine_start(ctas=64, idle_timeout_ms=8000)
                P.pack_fp8_write(ctx, x, stg, qp=qp, dst_mr=rmt, chunk_elems=chunk, signal_every=8)
                pr = min((P.pack_fp8_write(ctx, x, stg, qp=qp, dst_mr=rmt, chunk_elems=chunk, signal_every=8) for _ in range(3)),
                         key=lambda r: r.device_ns if r.ok else 1 << 62)      # best of 3 after one warm-up
                extras["fused_pack_fp8_write"] = {"elems": n_el, "ok": pr.ok, "device_us": round(pr.device_ns / 1e3, 1),
                                                  "source_bf16_gbps": round(pr.source_gbps, 1),
                                                  "wire_fp8_gbps": round(pr.payload_gbps, 1),
                                                  "frac_of_hbm_roofline": round(pr.source_gbps / R.fused_pack_roofline_gbps(), 3)}
        ctx.engine_start(ctas=args.engine_ctas, idle_timeout_ms=8000)
        sm = ops.rdma_stream(qp, W.OP_RDMA_WRITE, mr_slot[0], md_slot[0], 4096, iters=512, window=16, slot_stride=4096,
                             nslots=64, stream=stream)
        ops.rdma_stream(qp, W.OP_RDMA_WRITE, mr_slot[0], md_slot[0], 64, iters=32, window=1, stream=stream)
        lat = ops.rdma_stream(qp, W.OP_RDMA_WRITE, mr_slot[0], md_slot[0], 64, iters=128, window=1, stream=stream)
        extras["small_msg"] = {"4KiB_w16_us_per_msg": round(sm.us_per_msg, 2), "64B_latency_us": round(lat.us_per_msg, 2)}
        try:
            # perftest-style posting (--post_list / --cq-mod), clamped by the poster to what this QP's window allows
            bkw = dict(window=32, burst=8, signal_every=8, stream=stream)
            ops.rdma_stream(qp, W.OP_RDMA_WRITE, mr_slot[0], md_slot[0], 4096, iters=64, slot_stride=4096, nslots=64, **bkw)
            bm = ops.rdma_stream(qp, W.OP_RDMA_WRITE, mr_slot[0], md_slot[0], 4096, iters=2048, slot_stride=4096, nslots=64, **bkw)
            mb = ops.rdma_stream(qp, W.OP_RDMA_WRITE, mr_slot[0], md_slot[0], 1 << 20, iters=1024, slot_stride=1 << 20, nslots=64, **bkw)
            extras["small_msg"].update({"4KiB_burst8_us_per_msg": round(bm.us_per_msg, 3), "1MiB_burst8_gbps": round(mb.gbps, 1),
                                        "burst_ok": bool(bm.ok and mb.ok)})
        except Exception as e:  # extras never take the headline down
            extras["small_msg"]["burst_error"] = str(e)[:120]
    ctx.engine_stop()
    counters = qp.counters()
    torch.cuda.synchronize()
    barrier()

    # ---------------- max over ranks
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = t.tolist()
    ok = counters["n_err"] == 0 and counters["n_db_order_violations"] == 0
    if rank == 0:
        total_bytes = msg * args.steps * world
        value = total_bytes / (dev_ms_max * 1e-3) / 1e9
        e2e_value = total_bytes / (e2e_ms_max * 1e-3) / 1e9
        peaks = R.measured_peaks()
        out = {
            "metric": "rdma_write_gbps_gpu_hbm_device_timed",
            "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(dev_ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bytes (payload-agnostic; fused extra is bf16->fp8)", "data": "synthetic",
            "impl": "ours",
            "config": {"model": "gpu_initiated_rdma_write_loopback", "msg_bytes": msg, "global_batch": world,
                       "seq_len": msg, "parallelism": f"{world}x(GPU+own HCA), loopback per GPU (BASELINE config 5 shape)",
                       "wire": "softhca device engine over HBM (no /dev/infiniband in the container; CX-7 path gated off)",
                       "engine_ctas": args.engine_ctas, "steps_in_flight": depth, "poster": "sm_100a kernel: WQE + doorbell + device CQ poll",
                       "l2_policy": f"inputs larger than L2: {msg >> 20} MiB messages rotating over {nslots} buffers",
                       "timing": "CUDA events on the posting stream, max over ranks"},
            "roofline": {"bound_gbps_per_gpu": round(R.copy_roofline_gbps(peaks), 1),
                         "frac_of_measured_copy_peak": round(value / world / R.copy_roofline_gbps(peaks), 3),
                         "peaks_source": peaks.get("_source")},
            "e2e": {"value": round(e2e_value, 2), "unit": "GB/s", "h2d_bytes_per_step": msg, "d2h_bytes_per_step": 64,
                    "h2d_only_gbps": round(h2d_gbps, 1), "cpu_affinity": f"{len(cpus)} cpus local to the GPU" if cpus else "unbound",
                    "path": "pinned host -> cudaMemcpyAsync H2D -> GPU-posted RDMA write -> status words in mapped pinned memory"},
            "gpu_launches": args.steps, "gpu_launches_note": f"one poster kernel (one WQE) per step, {depth} steps in flight over {depth} QPs/streams; the DMA engine is one persistent kernel launched before the timed region",
            "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": clocks.get("reasons", []),
                       "samples": clocks.get("samples", 0)},
            "verified": ok, "extras": extras,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
