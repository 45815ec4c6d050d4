#!/bin/bash
export PYTHONPATH=/root/repo
export FK_LIB_PATH=/root/repo/frankenstein_b200/libfk_b200_ew.so
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "=== EW16"
timeout 200 python scripts/gpu_time_gemm.py 2>&1 | head -7 | cut -c1-260
echo "=== EW8"
FK_GEMM_EPI_WARPS=8 timeout 200 python scripts/gpu_time_gemm.py 2>&1 | head -7 | cut -c1-260
