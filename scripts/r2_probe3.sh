#!/bin/bash
# round-2 probe 3: full GPU test suite, bench with parity block, setup stage timing
cd /root/repo
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_p3_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_p3_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_p3_bench.json 2> gpurun_out/r2_p3_bench.err
echo "rc=$?" >> gpurun_out/r2_p3_bench.err
HDK_SETUP_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p3_timing.json 2> gpurun_out/r2_p3_timing.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_p3_ref.json 2> gpurun_out/r2_p3_ref.err
tail -5 gpurun_out/r2_p3_pytest.log
