#!/bin/bash
cd /root/repo
L=${PROBE_LIB:-/root/repo/hypredrive_b200/lib/libHYPREDRV.so}
run() { w=$1; shift; tag=$1; shift
  env HDK_LIB=$L HDK_IPC_TIMEOUT_S=20 "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$w --master-addr 127.0.0.1 --master-port 29611 tests/mp_gpu_check.py $ARGS > gpurun_out/r2_2gpu_$tag.log 2>&1
  echo "rc=$?" >> gpurun_out/r2_2gpu_$tag.log; grep -h "MPCHECK\|rc=" gpurun_out/r2_2gpu_$tag.log | cut -c1-260; }
ARGS="lap7 16 16 10";  run 2 sell HDK_REPLICATE_ROWS=40 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0
ARGS="convdif 16 8 6"; run 2 sellcd HDK_REPLICATE_ROWS=40 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0 MPCHECK_RAGGED=1
ARGS="lap7 16 14 7";   run 4 four HDK_REPLICATE_ROWS=60 MPCHECK_RAGGED=1 MPCHECK_SHARED_IPC=1 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0
ARGS="lap7 16 14 7";   run 2 nccl HDK_REPLICATE_ROWS=60 HDK_HALO_IPC=0 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0
i=0
for v in ${PROBE_VARIANTS:-default foldall rep100k}; do
  E=""
  [ $v = foldall ] && E="HDK_EXPORT_MAX_ROWS=1e12"
  [ $v = rep100k ] && E="HDK_REPLICATE_ROWS=100000"
  [ $v = packbig ] && E="HDK_EXPORT_MAX_ROWS=2000000"
  i=$((i+1))
  env HDK_LIB=$L HDK_TIMELINE=1 $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29740+i)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_2gpu_bench2_$v.json 2> gpurun_out/r2_2gpu_bench2_$v.err
  python - $v <<'P'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2_2gpu_bench2_%s.json'%v).read().strip().splitlines()[-1])
    print(v, 'value %.3e ms %.2f iters %d setup %.3f'%(d['value'],d['ms_per_step'],d['iterations'],d['setup_s']), 'e2e %.3e'%d['e2e']['value'])
    print('   ', {k:round(x['ms'],4) for k,x in d['kernels'].items()})
except Exception as e:
    print(v,'ERR',e)
P
  grep "hdk timeline rank 0" gpurun_out/r2_2gpu_bench2_$v.err | sed 's/\[hdk timeline rank 0\] //' | awk '{printf "%s | ", $0} END{print ""}' | cut -c1-2500
done
