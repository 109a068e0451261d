#!/bin/bash
# probe 8 (1 GPU): folded halo export through the peer-memory path with two ranks sharing the GPU
cd /root/repo
run() { w=$1; shift; tag=$1; shift
  /usr/bin/time -f "%e s" env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$w --master-addr 127.0.0.1 --master-port 29611 tests/mp_gpu_check.py $ARGS > gpurun_out/r2_p8_$tag.log 2>&1
  echo "rc=$?" >> gpurun_out/r2_p8_$tag.log; tail -n 3 gpurun_out/r2_p8_$tag.log | cut -c1-300; }
ARGS="lap7 16 16 10";  run 2 ipc_sell HDK_REPLICATE_ROWS=40 MPCHECK_SHARED_IPC=1 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0
ARGS="lap27 8 8 6";    run 2 ipc_sell27 HDK_REPLICATE_ROWS=30 MPCHECK_SHARED_IPC=1 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0
ARGS="convdif 16 8 6"; run 2 ipc_sellcd HDK_REPLICATE_ROWS=40 MPCHECK_SHARED_IPC=1 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0 MPCHECK_RAGGED=1
ARGS="lap7 16 14 7";   run 3 ipc_sell3 HDK_REPLICATE_ROWS=60 MPCHECK_SHARED_IPC=1 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0 MPCHECK_RAGGED=1
ARGS="lap7 16 16 10";  run 2 ipc_noexport HDK_REPLICATE_ROWS=40 MPCHECK_SHARED_IPC=1 HDK_SELL_MIN_ROWS=0 HDK_SELL_MIN_ROWS_DIST=0 HDK_HALO_EXPORT=0
ARGS="lap7 12 11 6";   run 2 ipc_stream HDK_REPLICATE_ROWS=40 MPCHECK_SHARED_IPC=1
