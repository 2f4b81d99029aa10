#!/bin/bash
# compute-sanitizer over the reduced shard: stream strategy, rescue hand-off forced by a small primary capacity
mkdir -p gpurun_out
export DBGPHMM_STRATEGY=stream
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # name, tool, extra env...
  name=$1; tool=$2; shift 2
  echo "=== $name ($tool) $*"
  env "$@" timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 7 python tools/san_case.py $ARGS > gpurun_out/san_$name.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|step |ok|Error|Invalid|Uninit|hazard" gpurun_out/san_$name.log | head -30
}
ARGS="60000 96 1500 2"
python tools/san_case.py $ARGS || exit 1
run mem_default memcheck A=1
run mem_cap48 memcheck DBGPHMM_SPARSE_CAP=48
run init_default initcheck A=1
run init_cap48 initcheck DBGPHMM_SPARSE_CAP=48
run race_cap48 racecheck DBGPHMM_SPARSE_CAP=48
run mem_group memcheck DBGPHMM_DENSE_GROUP=16
