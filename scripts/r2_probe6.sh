#!/bin/bash
cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p6_bench2.json 2> gpurun_out/r2_p6_bench2.err
echo "bench rc=$?"
HDK_MAILBOX=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p6_bench2_nomail.json 2> gpurun_out/r2_p6_bench2_nomail.err
python - <<'P'
import json
for f in ('r2_p6_bench2','r2_p6_bench2_nomail'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'value %.3e ms %.2f iters %d setup %.3f'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s']), 'e2e %.3e'%d['e2e']['value'], d['parity_companion'])
    print('   ', {k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
