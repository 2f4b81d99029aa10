"""Executed warp instructions and stall samples by SASS opcode, from `ncu --page source --csv` (plain or .gz).
usage: python tools/ncu_opcodes.py <source.csv[.gz]> [KERNEL substring]   -> one table per captured launch of that kernel (first one by default)"""
import collections, csv, gzip, re, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(gzip.open(path, "rt") if path.endswith(".gz") else open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for si, s in enumerate(starts):
    name = rows[s][1] if len(rows[s]) > 1 else ""
    if want and want not in name:
        continue
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    h = rows[s + 1]; ii = h.index("Instructions Executed"); sm = h.index("# Samples")
    cnt = collections.Counter(); st = collections.Counter()
    for r in rows[s + 2:end]:
        if len(r) <= max(ii, sm) or not r[0].startswith("0x"):
            continue
        ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", r[1].strip())
        op = ins.split()[0].rstrip(";")
        base = op.split(".")[0]
        if base == "IMAD":
            base = "IMAD." + ("MOV" if ".MOV" in op else "WIDE" if "WIDE" in op else "other")
        cnt[base] += int(r[ii] or 0); st[base] += int(r[sm] or 0)
    tot = sum(cnt.values()); ts = max(sum(st.values()), 1)
    print(f"== {name.split('(')[0]}: {tot} warp instructions executed, {ts} stall samples")
    for k, v in cnt.most_common(32):
        print(f"  {k:12s} {v / tot * 100:5.1f} % of instructions  {st[k] / ts * 100:5.1f} % of stall samples")
    break
