#!/bin/bash
cd /root/repo
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_lap7_256_n4.json 2> gpurun_out/r02_bench_lap7_256_n4.err
echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_lap7_256_n4.json').read().strip().splitlines()[-1]); print('N=4 value %.3e ms %.2f iters %d setup %.3f e2e %.3e'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s'],d['e2e']['value']))"
