#!/bin/bash
# one GPU: smoke, the -m gpu suite, the bench line as the driver runs it, and (optional) the CPU arm with full-length reads in its sample
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps ${1:-20} --warmup ${2:-5} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$3" = "cpufull" ]; then
  timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 --cpu-full-reads 2 > gpurun_out/ref_full.json 2> gpurun_out/ref_full.err; echo "ref full rc=$?"; cat gpurun_out/ref_full.json
fi
