mkdir -p gpurun_out
for c in 2 1; do echo "cluster=$c"; FK_GEMM_CLUSTER=$c timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -2; FK_GEMM_CLUSTER=$c timeout 200 python scripts/gpu_time_gemm.py 2>&1 | head -5; done 2>&1 | tee gpurun_out/r2i_gemm_cluster.log
echo "stages=all cluster=2"; FK_GEMM_STAGES=0 timeout 200 python scripts/gpu_time_gemm.py 2>&1 | head -4
