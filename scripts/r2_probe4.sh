#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_reference_callers.py tests/test_gpu_kernels.py -q > gpurun_out/r2_p4_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_p4_pytest.log
tail -40 gpurun_out/r2_p4_pytest.log
