#!/bin/bash
cd /root/repo
i=0
for v in default lb8 rep131k noexport; do
  E=""
  [ $v = lb8 ] && E="HDK_LIB=/root/repo/hypredrive_b200/lib/libHYPREDRV_lb8.so"
  [ $v = rep131k ] && E="HDK_REPLICATE_ROWS=131072"
  [ $v = noexport ] && E="HDK_HALO_EXPORT=0"
  i=$((i+1))
  env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29720+i)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p12_bench2_$v.json 2> gpurun_out/r2_p12_bench2_$v.err
done
python - <<'P'
import json
for v in ('default','lb8','rep131k','noexport'):
    f='r2_p12_bench2_'+v
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'value %.3e ms %.2f iters %d setup %.3f'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s']), 'e2e %.3e'%d['e2e']['value'])
    print('   ', {k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
