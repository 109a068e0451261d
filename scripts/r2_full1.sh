#!/bin/bash
# one-GPU round-2 pass: the whole GPU test suite, then the profiles, then BASELINE config C4
cd /root/repo
S=$SECONDS
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$? $((SECONDS-S))s"; tail -3 gpurun_out/r02_pytest_gpu.log
bash scripts/r2_profile.sh > gpurun_out/r02_profile.log 2>&1; echo "profile rc=$? $((SECONDS-S))s"
timeout 600 python bench.py --steps 5 --warmup 3 --config lap27_aniso_256 --no-cpu-baseline > gpurun_out/r02_bench_lap27_aniso_256_n1.json 2> gpurun_out/r02_bench_lap27_aniso_256_n1.err; echo "c4 rc=$? $((SECONDS-S))s"
tail -c 1500 gpurun_out/r02_bench_n1.json
