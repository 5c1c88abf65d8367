mkdir -p gpurun_out
for sh in "8192 8192 2048 chain 8192" "4096 4096 4096 chain 4096" "8192 4096 4096 chain 4096"; do timeout 300 python -m rocnrdma_b200.models.sendrecv_gemm $sh 2>&1 | tail -1 | tee -a gpurun_out/chain.jsonl; done
