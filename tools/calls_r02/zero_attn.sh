#!/bin/bash
# add_zero_attn on the GPU (key / value route with the batched-GEMM core and one more key in the column softmax)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pooler.py -m gpu -q -k "zero_attn or many_queries or north_star" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/zero_attn.log
