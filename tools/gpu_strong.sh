#!/bin/bash
mkdir -p gpurun_out
one() { tag=$1; cov=$2; shift 2; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline --no-e2e --reads-per-gpu 148 --coverage $cov > gpurun_out/st_$tag.json 2> gpurun_out/st_$tag.err; python - <<PY
import json
d=json.loads(open('gpurun_out/st_$tag.json').read().strip().splitlines()[-1])
print("$tag", "strong(%d reads) %.2f GCUPS (%.0f ms)" % (d["strong"]["reads_total"], d["strong"]["value"], d["strong"]["ms_per_step"]))
PY
}
one serial500 2.5 DBGPHMM_OVERLAP=0
one forced500 2.5 DBGPHMM_OVERLAP=1
one forced666 3.33 DBGPHMM_OVERLAP=1
one serial666 3.33 DBGPHMM_OVERLAP=0
one auto250 1.25 A=1
one forced1332 6.66 DBGPHMM_OVERLAP=1
one serial1332 6.66 DBGPHMM_OVERLAP=0
