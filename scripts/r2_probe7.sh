#!/bin/bash
cd /root/repo
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_p7_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_p7_pytest.log
tail -25 gpurun_out/r2_p7_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p7_bench.json 2> gpurun_out/r2_p7_bench.err
HDK_GRAPH_ROWS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p7_bench_nograph.json 2> gpurun_out/r2_p7_bench_nograph.err
python - <<'P'
import json
for f in ('r2_p7_bench','r2_p7_bench_nograph'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'value %.3e ms %.2f iters %d setup %.3f launches %d'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s'],d['gpu_launches']), 'e2e %.3e'%d['e2e']['value'])
    print('   ', {k:round(v['ms'],3) for k,v in d['kernels'].items()})
P
