#!/bin/bash
# reproduce the multi-GPU failure of round 1 with the driver's command (torchrun, 20 + 5 steps)
N=${1:-2}
mkdir -p gpurun_out
export BENCH_VERBOSE=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n$N.out 2> gpurun_out/n$N.err
rc=$?
echo "plain rc=$rc"; tail -3 gpurun_out/n$N.out; grep -v "^\[bench rank" gpurun_out/n$N.err | tail -30; grep "^\[bench rank" gpurun_out/n$N.err | tail -6
if [ $rc -ne 0 ]; then
  CUDA_LAUNCH_BLOCKING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n${N}b.out 2> gpurun_out/n${N}b.err
  echo "blocking rc=$?"; tail -3 gpurun_out/n${N}b.out; grep -v "^\[bench rank" gpurun_out/n${N}b.err | grep -i "FAILED\|error\|dbgphmm" | tail -20; grep "^\[bench rank" gpurun_out/n${N}b.err | tail -6
fi
nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv
