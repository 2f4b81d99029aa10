#!/bin/bash
# DESIGN.md 6b item 1, the first measurement of the next round: one GPU's read shard of BASELINE configs[4] (C5: N = 6.66 M nodes,
# 20 kbp reads) with the dense warm-up in groups of G reads over one pool of 2 G slabs (DBGPHMM_DENSE_GROUP=G), against the ungrouped
# run whose batch is limited to 296 reads by two 187 MB slabs per read.  Run under gpurun (one B200, ~12 min):
#     gpurun --timeout 1500 -- 'bash tools/measure_c5_groups.sh'
# Then read gpurun_out/c5_*.json: `value` (GCUPS) and `ms_per_step` per variant; the groups are worth switching on by default if
# the 1184-read step beats 4 x the 296-read step (the sparse phase then runs 8 jobs per SM instead of 2).
mkdir -p gpurun_out
C5="--genome-len 5000000 --read-len 20000 --steps 2 --warmup 2 --no-cpu-baseline"
set -x
timeout 600 python bench.py $C5 --reads-per-gpu 296 > gpurun_out/c5_ungrouped_296.json 2> gpurun_out/c5_ungrouped_296.err || tail -3 gpurun_out/c5_ungrouped_296.err
for G in 74 148 296; do
    DBGPHMM_DENSE_GROUP=$G timeout 900 python bench.py $C5 --reads-per-gpu 1184 > gpurun_out/c5_group${G}_1184.json 2> gpurun_out/c5_group${G}_1184.err \
        || tail -3 gpurun_out/c5_group${G}_1184.err
done
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"reads_per_gpu_per_step": [0-9]*' gpurun_out/c5_*.json
