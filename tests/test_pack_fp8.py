"""K3/K5: bf16 -> fp8 block-scaled pack (+ fused RDMA write) and unpack vs the PyTorch reference."""
import pytest
import torch

from rocnrdma_b200.ops import pack as P


def _payload(n, device="cpu", seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, generator=g) * torch.exp2(torch.randint(-12, 12, (n // 32,), generator=g).float()).repeat_interleave(32)
    x[:32] = 0.0                      # an all-zero block
    x[32:64] = 448.0                  # exactly representable maximum -> e = 0
    x[64:96] = 2.0 ** -20
    return x.to(torch.bfloat16).to(device)


# ------------------------------------------------------------------ CPU: the reference itself
def test_reference_scale_is_minimal_power_of_two():
    x = _payload(8192 * 4)
    rec = P.ref_pack_fp8(x, 8192)
    r = rec.reshape(4, P.record_bytes(8192))
    e = r[:, 8192:8192 + 256].to(torch.int32).reshape(-1) - 127
    amax = x.float().reshape(-1, 32).abs().amax(1)
    nz = amax > 0
    assert torch.all(amax[nz] / torch.exp2(e[nz].float()) <= 448.0)
    assert torch.all(amax[nz] / torch.exp2(e[nz].float() - 1) > 448.0), "scale is not the smallest admissible"
    assert e[1] == 0 and e[0] == -127


def test_reference_roundtrip_error_bound():
    x = _payload(8192 * 8, seed=3)
    rec = P.ref_pack_fp8(x, 8192 * 2)
    y = P.ref_unpack_fp8(rec, x.numel(), 8192 * 2).float()
    xf = x.float()
    amax = xf.reshape(-1, 32).abs().amax(1).repeat_interleave(32)
    # e4m3 has 3 mantissa bits: relative step 2^-3 at the top binade of the block -> half of that
    # (relative to the block's scaled maximum) bounds the absolute error, plus bf16 output rounding
    assert torch.all((y - xf).abs() <= amax * (2.0 ** -4) * 1.01 + 1e-30)


def test_record_layout_sizes():
    assert P.record_bytes(8192) == 8192 + 256
    assert P.record_bytes(1 << 20) == (1 << 20) + (1 << 15)
    assert P.staging_bytes(1 << 21, 1 << 20) == 2 * P.record_bytes(1 << 20)
    assert P.record_bytes(8192 * 3) % 64 == 0


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n_chunks,chunk", [(1, 8192), (3, 8192 * 4), (2, 1 << 20)])
def test_pack_only_matches_reference_bit_exact(ctx, n_chunks, chunk):
    x = _payload(n_chunks * chunk, "cuda:0", seed=n_chunks)
    staging = torch.zeros(P.staging_bytes(x.numel(), chunk), dtype=torch.uint8, device="cuda:0")
    smr = ctx.reg_mr(staging)
    torch.cuda.synchronize()
    P.pack_fp8_write(ctx, x, smr, qp=None, chunk_elems=chunk)
    ref = P.ref_pack_fp8(x, chunk)
    assert torch.equal(staging, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("with_imm", [False, True])
def test_fused_pack_and_rdma_write(ctx, with_imm):
    chunk, n_chunks = 1 << 18, 16
    x = _payload(chunk * n_chunks, "cuda:0", seed=9)
    nb = P.staging_bytes(x.numel(), chunk)
    staging = torch.zeros(nb, dtype=torch.uint8, device="cuda:0")
    remote = torch.zeros(nb, dtype=torch.uint8, device="cuda:0")
    out = torch.zeros_like(x)
    smr, rmr = ctx.reg_mr(staging), ctx.reg_mr(remote)
    cq_a, cq_b = ctx.create_cq(256), ctx.create_cq(256)
    qa = ctx.create_qp(cq_a, cq_a, 64, 64)
    qb = ctx.create_qp(cq_b, cq_b, 64, 64)
    qa.connect(qb)
    if with_imm:
        for _ in range(n_chunks):
            qb.post_recv(rmr, 0)
    torch.cuda.synchronize()
    ctx.engine_start(ctas=16, idle_timeout_ms=3000)
    try:
        if with_imm:
            uview, ustream = P.unpack_fp8(ctx, remote, out, chunk, qp=qb, sync=False, stream=ctx.aux_stream)
        r = P.pack_fp8_write(ctx, x, smr, qp=qa, dst_mr=rmr, chunk_elems=chunk, with_imm=with_imm, signal_every=4)
        if with_imm:
            ustream.synchronize()
    finally:
        ctx.engine_stop()
    assert r.ok and r.wqes == n_chunks, r
    ref = P.ref_pack_fp8(x, chunk)
    assert torch.equal(remote, ref), "records at the destination differ from the reference"
    c = qa.counters()
    assert c["n_wqe"] == n_chunks + 1 and c["n_err"] == 0 and c["n_db_order_violations"] == 0   # + flush NOP
    assert c["n_cqe"] == n_chunks // 4 + 1
    if with_imm:
        import ctypes as C
        w = (C.c_int64 * 8).from_buffer(uview)
        assert w[0] == 0 and w[3] == n_chunks, list(w)
        assert torch.equal(out, P.ref_unpack_fp8(ref, x.numel(), chunk))


@pytest.mark.gpu
def test_unpack_matches_reference(ctx):
    chunk = 8192 * 8
    x = _payload(chunk * 5, "cuda:0", seed=5)
    rec = P.ref_pack_fp8(x, chunk)
    out = torch.zeros_like(x)
    torch.cuda.synchronize()
    res = P.unpack_fp8(ctx, rec, out, chunk)
    assert res["status"] == "OK"
    assert torch.equal(out, P.ref_unpack_fp8(rec, x.numel(), chunk))
