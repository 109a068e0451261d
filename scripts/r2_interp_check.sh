#!/bin/bash
cd /root/repo
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py tests/test_gpu_large.py -m gpu -x -q -k "amg_setup or fixtures or vcycle_and_pcg or row_distributed or large or INTERP" 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_interp.json 2> gpurun_out/r02_bench_interp.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_interp.json').read().strip().splitlines()[-1]); print('value %.3e ms %.2f iters %d setup %.3f'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s']), d['setup_s_all'])"
HDK_SETUP_TIMING=1 timeout 200 python scripts/setup_probe.py 2>&1 | grep -i "interp\|level 0\|total" | head -12
