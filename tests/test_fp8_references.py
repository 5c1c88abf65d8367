"""CPU checks of the PyTorch references the GPU kernels are compared against (a wrong reference would make a wrong kernel pass):
block-scaled fp8 panel records (K4's fp8 epilogue), MX quantisation (K7's operands) and the chunk records of the pack (K3 / K5)."""
import torch

from rocnrdma_b200.ops import gemm as G
from rocnrdma_b200.ops import gemm_mx as MX
from rocnrdma_b200.ops import pack as P


def test_panel_records_round_trip_within_the_format_error_and_pad_ragged_m():
    torch.manual_seed(0)
    for M, N in ((256, 64), (300, 96), (130, 32)):
        x = torch.randn(M, N) * torch.exp2(torch.randint(-8, 8, (M, 1)).float())
        rec = G.ref_fp8_panels(x)
        panels = -(-M // 128)
        assert rec.numel() == panels * G.panel_record_bytes(N)
        back = G.dequant_fp8_panels(rec, M, N)
        blk = x.reshape(M, N // 32, 32).abs().amax(dim=2, keepdim=True).expand(-1, -1, 32).reshape(M, N)
        assert torch.all((back - x).abs() <= blk * 2.0 ** -4 * 1.01 + 1e-30)       # e4m3: 3 mantissa bits, scale = next power of two of amax / 448
        if M % 128:                                                                # rows past M quantise zeros: data bytes 0, scale = the zero-block exponent
            r = rec.reshape(panels, -1)
            tail_rows = r[-1, :128 * N].reshape(128, N)[M % 128:]
            assert int(tail_rows.max()) == 0


def test_scale_is_the_smallest_power_of_two_that_fits_448():
    # amax exactly 448 * 2^k must map to exponent k (not k + 1); just above it to k + 1
    for k in (-3, 0, 5):
        x = torch.zeros(1, 32); x[0, 0] = 448.0 * 2.0 ** k
        q, s = MX.quantize_mx(x)
        assert int(s[0, 0]) - 127 == k and q.view(torch.float8_e4m3fn).float()[0, 0] == 448.0
        x[0, 0] = 449.0 * 2.0 ** k
        q, s = MX.quantize_mx(x)
        assert int(s[0, 0]) - 127 == k + 1
    q, s = MX.quantize_mx(torch.zeros(2, 64))
    assert int(q.max()) == 0                                                       # all-zero blocks stay zero whatever their scale byte


def test_mx_quantisation_matches_the_panel_and_chunk_record_rules():
    """K3 (chunk records), K4 (panel records) and K7's operand quantiser must agree bit for bit: K7 consumes what the other two emit."""
    torch.manual_seed(1)
    M, K = 256, 128
    x = (torch.randn(M, K) * torch.exp2(torch.randint(-6, 7, (M, K // 32, 1)).float()).expand(M, K // 32, 32).reshape(M, K)).to(torch.bfloat16)
    q, s = MX.quantize_mx(x)
    rec = G.ref_fp8_panels(x.float()).reshape(M // 128, -1)
    assert torch.equal(rec[:, :128 * K].reshape(M, K), q) and torch.equal(rec[:, 128 * K:].reshape(M, K // 32), s)
    chunk = 128 * K                                                                # one chunk = one 128-row group
    crec = P.ref_pack_fp8(x.reshape(-1), chunk).reshape(M // 128, -1)
    assert torch.equal(crec[:, :chunk].reshape(M, K), q) and torch.equal(crec[:, chunk:chunk + chunk // 32].reshape(M, K // 32), s)
    back = P.ref_unpack_fp8(crec.reshape(-1), M * K, chunk).reshape(M, K)
    assert torch.equal(back.float(), MX.dequantize_mx(q, s).to(torch.bfloat16).float())
