"""Attribute ncu per-SASS-instruction counters to CUDA source lines (nvdisasm -g line table + ncu --page source --csv)."""
import csv, re, collections, sys
sass, srccsv, func, cufile = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
SORTCOL = 0 if (len(sys.argv) > 6 and sys.argv[6] == "inst") else 1
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and func in l][0]
addr2line = {}; cur = None
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//-----'): break
    m = re.search(r'line (\d+)', l)
    if '//## File' in l and m:
        cur = int(m.group(1)) if cufile in l else -abs(int(m.group(1))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/', l)
    if m and cur is not None: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
# the CSV holds one section per captured launch ("Kernel Name", <demangled name>): take the first one of the wanted kernel
# (KERNEL=<substring of the demangled name>, default: the first section)
import os
want = os.environ.get("KERNEL")
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
sec = None
for i in starts:
    if want is None or want in (rows[i][1] if len(rows[i]) > 1 else ''):
        sec = i; break
if sec is None: sys.exit(f"no section for kernel {want!r} in {srccsv}")
rows = rows[sec:]
h = rows[1]; ai = h.index('Address'); ii = h.index('Instructions Executed'); si = h.index('# Samples')
base = None; agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0]
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) < len(h): continue
    try: a = int(r[ai], 16)
    except Exception: continue
    if base is None: base = a
    ln = addr2line.get(a - base, 0)
    ie = int(r[ii] or 0); sm = int(r[si] or 0)
    agg[ln][0] += ie; agg[ln][1] += sm; agg[ln][2] += 1; tot[0] += ie; tot[1] += sm
src = open(cufile).read().split('\n')
print("total warp-inst", tot[0], "samples", tot[1])
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][SORTCOL])[:topn]:
    text = src[ln - 1].strip()[:100] if ln > 0 else f'(other file line {-ln})'
    print(f"{ln:5d} inst={v[0]/tot[0]*100:5.1f}% stall_samples={v[1]/max(tot[1],1)*100:5.1f}% nSASS={v[2]:3d} | {text}")
