#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -m gpu -x -q -k "fallback or timeline or peer_memory" 2>&1 | tail -5
