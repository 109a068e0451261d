#!/bin/bash
# round-2 probe 2: row-distributed setup -- 1 rank (forced), then 2/3/4 ranks sharing the GPU
cd /root/repo
run() { # world env... -- args
  w=$1; shift; tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$w --master-addr 127.0.0.1 --master-port 29611 tests/mp_gpu_check.py $ARGS > gpurun_out/r2_p2_$tag.log 2>&1
  echo "rc=$?" >> gpurun_out/r2_p2_$tag.log
}
ARGS="lap7 12 11 6";  run 1 w1_lap7 HDK_SETUP_DIST_FORCE=1 HDK_REPLICATE_ROWS=40
ARGS="lap7 12 11 6";  run 2 w2_lap7 HDK_REPLICATE_ROWS=40
ARGS="lap27 8 8 6";   run 2 w2_lap27 HDK_REPLICATE_ROWS=30
ARGS="convdif 16 8 6"; run 2 w2_convdif HDK_REPLICATE_ROWS=40 MPCHECK_RAGGED=1
ARGS="lap7 16 14 7";  run 3 w3_lap7 HDK_REPLICATE_ROWS=60 MPCHECK_RAGGED=1
ARGS="lap7 40 36 20"; run 2 w2_lap7_big HDK_REPLICATE_ROWS=2000
ARGS="lap7 20 18 9";  run 2 w2_lap7_deep HDK_REPLICATE_ROWS=10 MPCHECK_RAGGED=1
ARGS="lap7 12 11 6";  run 2 w2_lap7_ipc HDK_REPLICATE_ROWS=40 MPCHECK_SHARED_IPC=1
tail -n 3 gpurun_out/r2_p2_*.log
