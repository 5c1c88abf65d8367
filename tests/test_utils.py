"""Roofline arithmetic, clock-sample parsing, stream/result plumbing (CPU only)."""
import ctypes as C

from rocnrdma_b200.utils import roofline as R
from rocnrdma_b200.utils.clocks import ClockSampler


def test_roofline_uses_measured_peaks_and_payload_is_half_of_copy():
    p = R.measured_peaks()
    assert p["hbm_gbs"] > 1000 and p["_source"] in ("measured", "fallback")
    assert R.copy_roofline_gbps({"hbm_gbs": 6578.7}) == 6578.7 / 2
    # fused pack: the algorithmic bytes -- 2 B read + (1 + 1/32) B written per 2 B of source (no fraction above 1 any more)
    src = R.fused_pack_roofline_gbps({"hbm_gbs": 6578.7})
    assert abs(src - 6578.7 * 2 / (2 + 33 / 32)) < 1e-6
    assert R.fused_pack_roofline_gbps({"hbm_gbs": 6578.7}, wire_gbs=50.0) == 50.0 * 2 / (33 / 32)   # NIC-bound on a real wire
    assert R.fraction(50.0, 100.0) == 0.5


def test_clock_sampler_summary_flags_throttle_reasons():
    s = ClockSampler()
    s.rows = [dict(index=0, sm=1965.0, sm_max=1965.0, power=400.0, active="0x0", hw_slowdown="Not Active", hw_thermal="Not Active",
                   sw_thermal="Not Active", sw_power="Active"),
              dict(index=0, sm=1300.0, sm_max=1965.0, power=990.0, active="0x4", hw_slowdown="Not Active", hw_thermal="Not Active",
                   sw_thermal="Not Active", sw_power="Active"),
              dict(index=0, sm=1400.0, sm_max=1965.0, power=980.0, active="0x4", hw_slowdown="Active", hw_thermal="Not Active",
                   sw_thermal="Not Active", sw_power="Not Active")]
    out = s.summary()
    assert out["sm_mhz"] == 1400.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["hw_slowdown", "sw_power_cap"]
    assert ClockSampler().summary()["samples"] == 0


def test_stream_result_arithmetic():
    from rocnrdma_b200.ops.rdma import StreamResult, parse_stream_out
    r = StreamResult(status=["OK", "OK"], t_start_ns=[1000, 1100], t_end_ns=[3000, 3100], done=[4, 4], bytes_per_msg=1 << 20)
    assert r.ok and r.device_ns == 2100 and abs(r.gbps - 8 * (1 << 20) / 2100) < 1e-9 and abs(r.us_per_msg - 2.1 / 4) < 1e-9
    buf = (C.c_uint8 * 128)()
    words = (C.c_int64 * 16).from_buffer(buf)
    words[0], words[1], words[2], words[3] = -1, 5, 9, 3         # CTA 0 timed out
    words[8], words[9], words[10], words[11] = 0, 6, 10, 7
    p = parse_stream_out(buf, 2, 64)
    assert p.status == ["TIMEOUT", "OK"] and not p.ok and p.done == [3, 7] and p.device_ns == 5


def test_pack_and_gemm_record_geometry():
    from rocnrdma_b200.ops import gemm, pack
    assert pack.record_bytes(1 << 22) == (1 << 22) + (1 << 17)
    assert gemm.panel_record_bytes(8192) == 128 * 8192 * 33 // 32


def test_stream_posting_parameters_always_allow_progress():
    """Device poster: burst / cq-moderation are clamped so the window wait always has a signaled WQE to wait for."""
    import ctypes as C
    from rocnrdma_b200 import _native as N
    lib = N.load()
    for window in (0, 1, 2, 3, 8, 16, 33, 128):
        for burst in (0, 1, 7, 16, 32, 500):
            for sig in (0, 1, 4, 16, 100, 1000):
                b, s = C.c_uint32(burst), C.c_uint32(sig)
                lib.rn_stream_clamp(window, C.byref(b), C.byref(s))
                assert 1 <= b.value <= 32 and s.value >= 1
                if window:
                    assert b.value <= (window + 1) // 2
                    # newest signaled WQE is at most s-1 behind the newest posted; the wait target is window-b behind
                    assert s.value - 1 <= window - b.value
                else:
                    assert b.value == min(max(burst, 1), 32) and s.value == max(sig, 1)


def test_fp8_panel_reference_roundtrip_and_record_geometry():
    """The PyTorch reference of the GEMM's fp8 epilogue (what the GPU test compares the kernel with):
    record geometry, UE8M0 scales that are the smallest power of two keeping the block within e4m3 range,
    and a dequantised error bounded by the e4m3 step at each block's scale."""
    import torch
    from rocnrdma_b200.ops import gemm as G
    torch.manual_seed(3)
    M, Nn = 256, 512
    c = torch.randn(M, Nn) * torch.logspace(-6, 6, Nn // 32).repeat_interleave(32)[None, :]      # 12 decades of block magnitudes
    c[5, 64:96] = 0.0                                                                              # an all-zero block
    rec = G.ref_fp8_panels(c)
    assert rec.dtype == torch.uint8 and rec.numel() == (M // 128) * G.panel_record_bytes(Nn)
    assert G.panel_record_bytes(Nn) == 128 * Nn + 128 * (Nn // 32)
    back = G.dequant_fp8_panels(rec, M, Nn)
    blocks = c.reshape(M, Nn // 32, 32)
    amax = blocks.abs().amax(dim=2)
    r = rec.reshape(M // 128, G.panel_record_bytes(Nn))
    e = r[:, 128 * Nn:].to(torch.int32).reshape(M, Nn // 32) - 127
    scale = torch.pow(2.0, e.double())
    nz = amax > 0
    assert bool(((amax.double() / scale)[nz] <= 448.0).all())                 # fits e4m3 after scaling ...
    assert bool(((amax.double() / (scale / 2))[nz] > 448.0).all())            # ... and no smaller power of two would
    err = (back.reshape(M, Nn // 32, 32) - blocks).abs().double()
    assert bool((err <= (scale * 32.0)[..., None] * 0.5 + 1e-30).all())       # half of e4m3's largest step (2^5 at 2^8) times the scale
    assert bool((back.reshape(M, Nn // 32, 32)[5, 2] == 0).all())
