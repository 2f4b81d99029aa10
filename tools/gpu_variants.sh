#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  echo "== variant $v"
  DBGPHMM_LIB_PATH=dbgphmm_b200/lib/variants/$v/libdbgphmm_b200.so timeout 600 python tools/profile_step.py --reads 1184 --read-len 1500 --reps 3 2>&1 | grep -E "^rep|Error|error" | tail -2
done
