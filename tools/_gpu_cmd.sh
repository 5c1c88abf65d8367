timeout 600 python -m pytest tests/test_gemm_mx.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12
RN_MX_CTA_GROUP=2 timeout 300 python tools/mx_bench.py 2>&1 | tail -3
python - <<'PY'
import torch, rocnrdma_b200 as rn
from rocnrdma_b200 import ops
from rocnrdma_b200.ops import gemm_mx as MX
ctx = rn.Context(0, wire="softhca")
for (M,N,K) in ((4096,4096,4096),(8192,8192,8192),(8192,8192,2048)):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    c = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    (aq, as_), (bq, bs) = MX.quantize_mx(a), MX.quantize_mx(b)
    oa, ob = MX.MxOperand.from_tensors(aq, as_), MX.MxOperand.from_tensors(bq, bs)
    for rep in range(3):
        r = ops.gemm_mxfp8(ctx, oa, ob, c, cta_group=2)
    print(M,N,K, f'{r.tflops:.0f} TF in-kernel, {r.device_ns/1e3:.1f} us', r.issuer_cycles, 'k-blocks/tile', K//128)
ctx.close()
PY
