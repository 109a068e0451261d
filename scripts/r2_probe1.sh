#!/bin/bash
# round-2 probe 1: does a 2-rank NCCL job run on ONE GPU (NCCL_HOSTID per rank, loopback sockets)?
cd /root/repo
nvidia-smi -L > gpurun_out/r2_p1_smi.txt 2>&1
ip addr > gpurun_out/r2_p1_ip.txt 2>&1
nproc > gpurun_out/r2_p1_nproc.txt
for args in "lap7 12 11 6" "convdif 16 8 6"; do
NCCL_DEBUG=WARN HDK_REPLICATE_ROWS=40 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/mp_gpu_check.py $args > gpurun_out/r2_p1_mp_${args// /_}.log 2>&1
echo "rc=$?" >> gpurun_out/r2_p1_mp_${args// /_}.log
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_p1_pytest.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_p1_bench.json 2> gpurun_out/r2_p1_bench.err
