#!/bin/bash
# the headline configuration (N = 32768, 64 + 16 pairs) against the fp32 eager restatement on the same GPU
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k headline --durations=3 2>&1 | tail -25 | cut -c1-400 | tee gpurun_out/fullsize_headline.log
