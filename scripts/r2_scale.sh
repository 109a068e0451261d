#!/bin/bash
# round-2 scaling runs on ONE 8-GPU box: headline weak scaling at N = 8 and the BASELINE configs C3 / C5
cd /root/repo
run() { # n config tag extra-args...
  n=$1; cfg=$2; tag=$3; shift 3
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
      bench.py --gpus $n --steps 5 --warmup 3 --config $cfg "$@" > gpurun_out/r02_bench_${tag}.json 2> gpurun_out/r02_bench_${tag}.err
  echo "$tag rc=$? t=$SECONDS"
}
run 8 lap7_256 lap7_256_n8
run 8 lap7_512_strong lap7_512_strong_n8 --no-cpu-baseline
run 8 convdif_gmres_256 convdif_gmres_256_n8 --no-cpu-baseline
[ -n "$SCALE_N4" ] && run 4 lap7_256 lap7_256_n4 --no-cpu-baseline
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*_n[48].json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f.split('r02_bench_')[1], 'N=%d value %.3e ms %.2f iters %d setup %.3f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['iterations'],d['setup_s']), 'e2e %.3e'%d['e2e']['value'], 'companion', (d.get('parity_companion') or {}).get('ok'), 'vcycle %.3f'%d['kernels']['vcycle']['ms'])
P
