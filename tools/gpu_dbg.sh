#!/bin/bash
mkdir -p gpurun_out
BENCH_VERBOSE=1 DBGPHMM_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-extras --no-cpu-baseline --coverage 10 > gpurun_out/dbg.json 2> gpurun_out/dbg.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/dbg.json').read().strip().splitlines()[-1])
print("value %.2f ms %.1f e2e %.2f strong %.2f (%.0f ms, %d reads)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["strong"]["value"], d["strong"]["ms_per_step"], d["strong"]["reads_total"]))
PY
grep -n "bench rank\|groups of\|splitting\|exhausted\|trace\] batch\|alloc_pool\|release\|fwd setup\|run_forward \|run_backward \|recompute\|products" gpurun_out/dbg.err | tail -90
