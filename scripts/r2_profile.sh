#!/bin/bash
# round-2 profiles (ONE GPU): bench line, then (each after its own command has exited 0 without ncu)
#   1. ncu launch list of one timed bench step,
#   2. ncu --set full of the sliced-ELL kernel (fine level, levels 1-2, fused variants),
#   3. ncu launch list + selected metrics of one AMG setup.
cd /root/repo
export HDK_HALO_IPC=0
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || exit 1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
HDK_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 2000 --csv \
   --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu1.log 2>&1
HDK_PROFILE_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spmv_sell -c 16 \
   -f -o gpurun_out/r02_spmv_sell python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu2.log 2>&1
timeout 300 python scripts/setup_probe.py > gpurun_out/r02_setup_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r02_setup_raw.csv python scripts/setup_probe.py > gpurun_out/r02_ncu3.log 2>&1
ls -la gpurun_out | grep r02_
