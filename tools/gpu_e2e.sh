#!/bin/bash
mkdir -p gpurun_out
BENCH_VERBOSE=1 DBGPHMM_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_t.json').read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "strong", d.get("strong",{}).get("value"))
PY
grep -n "bench rank\|groups of\|splitting\|exhausted\|trace\] batch\|release\|alloc_pool" gpurun_out/bench_t.err | tail -60
