#!/bin/bash
# Round measurement on one B200: bench line, ncu launch list, ncu --set full captures of the dense pair kernels and of k_sparse.
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
set -x
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err || { tail -5 gpurun_out/bench_1gpu.err; exit 1; }
cat gpurun_out/bench_1gpu.json
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --reads-per-gpu 148 > gpurun_out/b148.log 2>&1 || exit 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --reads-per-gpu 148 > gpurun_out/ncu_list.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b1184.log 2>&1 || exit 3
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_dense_(fwd|bwd)2' -s 18 -c 4 -f -o gpurun_out/dense_pair \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_dense.log 2>&1
timeout 300 python tools/profile_step.py --reads 1184 --read-len 1500 --reps 1 > gpurun_out/p1500.log 2>&1 || exit 4
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_sparse' -c 2 -f -o gpurun_out/sparse \
    python tools/profile_step.py --reads 1184 --read-len 1500 --reps 1 > gpurun_out/ncu_sparse.log 2>&1
ls -la gpurun_out
