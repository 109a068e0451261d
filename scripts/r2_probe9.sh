#!/bin/bash
# probe 9 (1 GPU): everything since the last green run -- full GPU suite, bench, variants
cd /root/repo
S=$SECONDS
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_p9_pytest.log 2>&1
echo "rc=$? elapsed=$((SECONDS-S))s" >> gpurun_out/r2_p9_pytest.log
tail -30 gpurun_out/r2_p9_pytest.log | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_p9_bench.json 2> gpurun_out/r2_p9_bench.err
HDK_LIB=/root/repo/hypredrive_b200/lib/libHYPREDRV_xhint.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p9_bench_xhint.json 2> gpurun_out/r2_p9_bench_xhint.err
HDK_RAP_SINGLE=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p9_bench_rap2.json 2> gpurun_out/r2_p9_bench_rap2.err
HDK_SETUP_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p9_timing.json 2> gpurun_out/r2_p9_timing.err
python - <<'P'
import json
for f in ('r2_p9_bench','r2_p9_bench_xhint','r2_p9_bench_rap2'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'value %.3e ms %.2f iters %d setup %.3f launches %d'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s'],d['gpu_launches']), 'e2e %.3e'%d['e2e']['value'], 'parity', (d.get('parity') or {}).get('ok'))
    print('   ', {k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
