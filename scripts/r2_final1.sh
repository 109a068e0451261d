#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke(), both bench arms; then C5 / C3 on one GPU
cd /root/repo
S=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$? $((SECONDS-S))s"; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err; echo "bench rc=$? $((SECONDS-S))s"
timeout 300 python bench.py --steps 5 --warmup 3 --config convdif_gmres_256 --no-cpu-baseline > gpurun_out/r02_bench_convdif_gmres_256_n1.json 2> gpurun_out/r02_bench_convdif_gmres_256_n1.err; echo "c5 rc=$? $((SECONDS-S))s"
timeout 400 python bench.py --steps 3 --warmup 3 --config lap7_512_strong --no-cpu-baseline > gpurun_out/r02_bench_lap7_512_strong_n1.json 2> gpurun_out/r02_bench_lap7_512_strong_n1.err; echo "c3 rc=$? $((SECONDS-S))s"
python - <<'P'
import json
for f in ('r02_bench_default_final','r02_bench_convdif_gmres_256_n1','r02_bench_lap7_512_strong_n1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value %.3e ms %.2f iters %d setup %.3f e2e %.3e'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s'],d['e2e']['value']), 'parity', (d.get('parity') or {}).get('ok'), 'roof %.3f'%d['roofline']['frac'], 'launches', d['gpu_launches'])
    except Exception as e:
        print(f,'ERR',e)
P
