#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-strong --no-extras"
timeout 300 $B > gpurun_out/b_plain.log 2>&1 || { echo "bench failed"; exit 3; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
DBGPHMM_VERIFY=1 timeout 900 python tools/soak.py 30 1332 1000 > gpurun_out/soak_v.log 2>&1; echo "soak verify rc=$?"; tail -2 gpurun_out/soak_v.log
timeout 900 python tools/soak.py 60 1332 3000 > gpurun_out/soak_p.log 2>&1; echo "soak plain rc=$?"; tail -2 gpurun_out/soak_p.log
