#!/bin/bash
# probe 10 (2 GPUs): folded exports + mailbox on real peers
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r2_p10_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_p10_pytest.log; tail -5 gpurun_out/r2_p10_pytest.log | cut -c1-300
for v in default noexport; do
  E=""; [ $v = noexport ] && E="HDK_HALO_EXPORT=0"
  env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_p10_bench2_$v.json 2> gpurun_out/r2_p10_bench2_$v.err
done
python - <<'P'
import json
for f in ('r2_p10_bench2_default','r2_p10_bench2_noexport'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'value %.3e ms %.2f iters %d setup %.3f'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s']), 'e2e %.3e'%d['e2e']['value'], d['parity_companion'])
    print('   ', {k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
