#!/bin/bash
# ncu --set full of k_mapx (the C4 production kernel) inside tools/bench_c4.py ; the raw page comes back as CSV
mkdir -p gpurun_out
timeout 600 python tools/bench_c4.py --steps 2 > gpurun_out/c4.json 2> gpurun_out/c4.err || { tail -5 gpurun_out/c4.err; exit 3; }
cat gpurun_out/c4.json
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_mapx' -s 1 -c 1 -f -o /tmp/mapx python tools/bench_c4.py --steps 1 > gpurun_out/ncu_mapx.log 2>&1
echo "ncu mapx rc=$?"
ncu -i /tmp/mapx.ncu-rep --page raw --csv > gpurun_out/raw_mapx.csv 2>/dev/null
ncu -i /tmp/mapx.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/src_mapx.csv.gz
ls -la gpurun_out/raw_mapx.csv gpurun_out/src_mapx.csv.gz
