#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "c4_full_size" > gpurun_out/pytest_c4.log 2>&1; echo "c4 test rc=$?"; tail -3 gpurun_out/pytest_c4.log
DBGPHMM_TRACE=1 timeout 1500 python bench.py --genome-len 5000000 --read-len 20000 --reads-per-gpu ${1:-1184} --steps 2 --warmup 1 --no-cpu-baseline --no-strong --no-extras --no-e2e > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "c5 rc=$?"; cat gpurun_out/bench_c5.json; grep -E "groups of|carried on|splitting|exhausted|FAILED" gpurun_out/bench_c5.err | sort | uniq -c | head; grep "trace\] batch\|run_forward \|run_backward \|recompute" gpurun_out/bench_c5.err | tail -8
