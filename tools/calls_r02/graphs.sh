#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graphs.py -m gpu -q -x 2>&1 | tail -30 | cut -c1-300 | tee gpurun_out/graphs.log
