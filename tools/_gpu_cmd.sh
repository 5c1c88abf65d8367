timeout 600 python -m pytest tests/test_gemm_mx.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
