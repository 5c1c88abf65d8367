timeout 600 python -m pytest tests/test_gemm_mx.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
RN_MX_CTA_GROUP=2 timeout 300 python tools/mx_bench.py 2>&1 | tail -3
