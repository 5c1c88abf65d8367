#!/bin/bash
# compute-sanitizer over the product kernels on one GPU (memcheck everywhere, racecheck on the kernels that use shared
# memory / mbarriers).  The persistent engine runs under the tool too, so timeouts are stretched via pytest's own bounds.
# Output: gpurun_out/sanitizer_{memcheck,racecheck}.txt (copy the summaries to profiles/).
set -u
mkdir -p gpurun_out
SEL_MEM='test_gpu_posted_write_sizes or test_gpu_posted_read_and_unaligned or test_many_ctas_one_sq_stress or test_burst_stream_many_qps_soak or test_bad_rkey or test_fused_pack_and_rdma_write or test_pack_only_matches or test_unpack_matches or test_gemm_cta_pair_kernel_matches_fp32_reference or test_gemm_cta_pair_send_panels or test_gemm_fp8_epilogue or test_gpu_initiated_write_into_mlx5dv_queues or test_fused_pack_posts_to_the_nic or test_gemm_ragged or test_gemm_wide_pair_kernel_matches_fp32_reference or test_gemm_wide_send_panels or test_mxfp8 or test_counters_report or test_host_post_beyond or test_unpolled_cq or test_shared_send_and_receive or test_reset_forgets'
SEL_RACE='test_gpu_posted_write_sizes or test_many_ctas_one_sq_stress or test_pack_only_matches or test_fused_pack_and_rdma_write or test_gemm_cta_pair_kernel_matches_fp32_reference or test_gemm_compute_only_matches_fp32_reference or test_gemm_cta_pair_send_panels or test_gemm_wide_pair_kernel_matches_fp32_reference or test_gemm_ragged_fp8_records or test_mxfp8_gemm_matches'
summ() {   # full log -> short evidence file: every distinct finding once (no host backtraces), the pytest line, the tool's summary
  grep -E "^========= (Invalid|Error|Warning|Program hit|Race reported|    and |    at |.* bytes|     Address|RACECHECK|ERROR SUMMARY|COMPUTE-SANITIZER)|passed|failed" "$1" | grep -v "Host Frame" | awk '!seen[$0]++' | head -150
}
timeout 1500 compute-sanitizer --tool memcheck --report-api-errors no --launch-timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "$SEL_MEM" > /tmp/memcheck_full.txt 2>&1
echo "memcheck rc=$?" >> /tmp/memcheck_full.txt
summ /tmp/memcheck_full.txt > gpurun_out/sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 120 python -m pytest tests -m gpu -q -rf -p no:cacheprovider -k "$SEL_RACE" > /tmp/racecheck_full.txt 2>&1
echo "racecheck rc=$?" >> /tmp/racecheck_full.txt
summ /tmp/racecheck_full.txt > gpurun_out/sanitizer_racecheck.txt
grep -E "^(FAILED|ERROR) " /tmp/racecheck_full.txt | cut -c1-300 >> gpurun_out/sanitizer_racecheck.txt
head -c 3000000 /tmp/memcheck_full.txt > gpurun_out/sanitizer_memcheck_full.txt
tail -n 12 gpurun_out/sanitizer_memcheck.txt; tail -n 12 gpurun_out/sanitizer_racecheck.txt
