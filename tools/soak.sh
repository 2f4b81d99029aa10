#!/bin/bash
mkdir -p gpurun_out
DBGPHMM_VERIFY=1 timeout 900 python tools/soak.py ${1:-60} 1184 1000 > gpurun_out/soak_v.log 2>&1; echo "verify rc=$?"; grep -v "^step .* ok" gpurun_out/soak_v.log | tail -20
timeout 900 python tools/soak.py ${2:-100} 1184 2000 > gpurun_out/soak_p.log 2>&1; echo "plain rc=$?"; grep -v "^step .* ok" gpurun_out/soak_p.log | tail -20; tail -2 gpurun_out/soak_p.log
