#!/bin/bash
# 2 B200s: the multi-rank NCCL parity tests and the column-side layout test on the final code of round 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py "tests/test_gpu_infonce.py::test_stored_exponential_kernels_match_recompute_kernels" -m gpu -q > gpurun_out/pytest_multirank_2gpu_r02h.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multirank_2gpu_r02h.log; tail -5 gpurun_out/pytest_multirank_2gpu_r02h.log
