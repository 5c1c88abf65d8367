"""Workloads ("models" of this framework are traffic patterns, not networks -- the reference has neither):

  loopback        BASELINE configs 2 / 3 / 5: one GPU driving its own HCA (write / read / fused pack)
  sendrecv_gemm   BASELINE config 4: GEMM on GPU0 producing panels that land on GPU1
"""
from . import loopback, sendrecv_gemm  # noqa: F401
