#!/bin/bash
# Round measurement on one B200: full-size parity test, ncu --set full captures (with source) of the dense pair kernels and of k_sparse.
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
# gpurun brings back at most 64 MiB: the reports stay on the box, their raw / source pages come back as (gzipped) CSV.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-strong --no-extras"
if [ "$1" != "noparity" ]; then
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q -s > gpurun_out/pytest_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -12 gpurun_out/pytest_fullsize.log
fi
timeout 300 $B > gpurun_out/b1184.log 2>&1 || { echo "bench failed"; tail -5 gpurun_out/b1184.log; exit 3; }
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_dense_(fwd|bwd)2' -s 18 -c 4 -f -o /tmp/dense_pair $B > gpurun_out/ncu_dense.log 2>&1
echo "ncu dense rc=$?"
ncu -i /tmp/dense_pair.ncu-rep --page raw --csv > gpurun_out/raw_dense_pair.csv 2>/dev/null
ncu -i /tmp/dense_pair.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/src_dense_pair.csv.gz
timeout 300 python tools/profile_step.py --reads 1184 --read-len 1500 --reps 1 > gpurun_out/p1500.log 2>&1 || { echo "profile_step failed"; tail -5 gpurun_out/p1500.log; exit 4; }
cat gpurun_out/p1500.log | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_sparse' -c 4 -f -o /tmp/sparse python tools/profile_step.py --reads 1184 --read-len 1500 --reps 1 > gpurun_out/ncu_sparse.log 2>&1
echo "ncu sparse rc=$?"
ncu -i /tmp/sparse.ncu-rep --page raw --csv > gpurun_out/raw_sparse.csv 2>/dev/null
ncu -i /tmp/sparse.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/src_sparse.csv.gz
# sparse capacity / residency experiment: jobs per SM by shared memory, filled waves through the grouped dense warm-up
for cfg in "128 1184 0" "112 1480 740" "96 1480 740" "96 1628 814" "80 1776 888"; do
  set -- $cfg
  echo "== cap $1 reads $2 group $3"
  DBGPHMM_TRACE=1 DBGPHMM_SPARSE_CAP=$1 DBGPHMM_DENSE_GROUP=$3 timeout 600 python tools/profile_step.py --reads $2 --read-len 10000 --reps 2 2>&1 | grep -E "^rep|carried on|need a larger|FAILED|Error" | tail -8
done
ls -la gpurun_out | tail -12; du -sh gpurun_out
