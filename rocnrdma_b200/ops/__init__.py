"""Python entry points for the sm_100a kernels (K1-K6 of SURVEY.md section 2.3)."""
from .rdma import rdma_stream, StreamResult, fill_random, fill_bf16, checksum, compare, l2_flush  # noqa: F401
