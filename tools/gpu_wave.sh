#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" DBGPHMM_TRACE=1 timeout 600 python tools/profile_step.py --reads $R --read-len 10000 --reps 2 2>&1 | grep -E "^rep|carried on|groups of|need a larger|FAILED|rror" | tail -4; }
R=1332 run DBGPHMM_SPARSE_RESCUE=192 DBGPHMM_SPARSE_MARGIN=0
R=1332 run DBGPHMM_SPARSE_RESCUE=176 DBGPHMM_SPARSE_MARGIN=0
R=1184 run A=1
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -s -k "c4_full_size" 2>&1 | grep -v "^$" | tail -12
