#!/bin/bash
# N GPUs: the NCCL parity test, then the driver's bench command
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nccl_gpu.py -m gpu -x -q > gpurun_out/pytest_nccl.log 2>&1; echo "nccl test rc=$?"; tail -4 gpurun_out/pytest_nccl.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${2:-20} --warmup ${3:-5} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.json; grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/bench_n$N.err | tail -15
