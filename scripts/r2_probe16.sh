#!/bin/bash
# per-operation timeline of one solve at N=1 and N=2 (same build)
cd /root/repo
L=${PROBE_LIB:-/root/repo/hypredrive_b200/lib/libHYPREDRV.so}
HDK_LIB=$L HDK_TIMELINE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p16_n1.json 2> gpurun_out/r2_p16_n1.err
HDK_LIB=$L HDK_TIMELINE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29750 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p16_n2.json 2> gpurun_out/r2_p16_n2.err
grep "hdk timeline" gpurun_out/r2_p16_n1.err | head -40
grep "hdk timeline rank 0" gpurun_out/r2_p16_n2.err | head -40
