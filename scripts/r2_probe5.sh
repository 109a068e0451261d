#!/bin/bash
# round-2 probe 5 (2 GPUs): multi-rank tests on real peers (IPC halo, mailbox reductions), bench N=2
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r2_p5_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_p5_pytest.log
tail -15 gpurun_out/r2_p5_pytest.log
HDK_SETUP_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_p5_bench2.json 2> gpurun_out/r2_p5_bench2.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_p5_bench2.json
HDK_MAILBOX=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p5_bench2_nomail.json 2> gpurun_out/r2_p5_bench2_nomail.err
HDK_SETUP_REPLICATED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p5_bench2_repl.json 2> gpurun_out/r2_p5_bench2_repl.err
