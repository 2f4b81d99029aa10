"""Static instruction mix of the kernels of one object file, from `cuobjdump -sass` (no GPU needed).

    python tools/sass_mix.py dbgphmm_b200/lib/dense.o k_dense_fwd2 k_dense_bwd2 > profiles/rN_sass_mix.txt

Counts SASS instructions per kernel by class (FP64 arithmetic, shared-memory, global-memory, shuffles, integer / logic, predicate
and branch, ...).  It is the static code, not the executed stream: read it next to the executed counts of the ncu capture
(profiles/*_ncu.txt).  Useful before spending GPU time: what share of a hot kernel's code is arithmetic, where the spills sit
(STL / LDL), whether bulk copies compiled to LDGSTS (cp.async)."""
import collections
import re
import subprocess
import sys

CLASSES = [
    ("fp64", r"^(DADD|DMUL|DFMA|DSETP|DMNMX|MUFU\.RCP64H|F2F\.F64|I2F\.F64|F2I\.\w*F64|D2I|I2D)"),
    ("shared ld/st", r"^(LDS|STS|LDSM|ATOMS)"),
    ("cp.async (LDGSTS) + fences", r"^(LDGSTS|LDGDEPBAR|DEPBAR|MEMBAR|FENCE|CCTL|ERRBAR)"),
    ("global ld/st", r"^(LDG|STG|LD\.|ST\.|ATOMG|RED|ATOM)"),
    ("local (spill) ld/st", r"^(LDL|STL)"),
    ("constant / uniform loads", r"^(LDC|ULDC|S2R|S2UR|CS2R)"),
    ("shuffle / vote / match", r"^(SHFL|VOTE|VOTEU|MATCH|REDUX)"),
    ("barrier / sync", r"^(BAR|WARPSYNC|BSSY|BSYNC|NANOSLEEP|YIELD)"),
    ("branch / exit", r"^(BRA|BRX|JMP|CALL|RET|EXIT|BREAK|BMOV)"),
    ("predicate / select", r"^(ISETP|PLOP3|P2R|R2P|SEL|FSEL|PSETP|UISETP|UPLOP3|USEL)"),
    ("integer / logic / move", r"^(IADD|IADD3|IMAD|LOP3|LOP|SHF|SHL|SHR|LEA|MOV|PRMT|IABS|IMNMX|FLO|POPC|BREV|BMSK|SGXT|I2I|UIADD3|UIMAD|ULOP3|USHF|ULEA|UMOV|UFLO|UPOPC|R2UR|VIADD|VIMNMX|UPRMT|I2IP|LEPC|UBMSK|USGXT|UIMNMX|VIADDMNMX)"),
    ("fp32 / conversions", r"^(FADD|FMUL|FFMA|FSETP|FMNMX|MUFU|F2F|I2F|F2I|FRND|HADD|HMUL|HFMA)"),
]


def main():
    obj, want = sys.argv[1], sys.argv[2:]
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    parts = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    print(f"# static SASS instruction mix, {obj} (cuobjdump -sass; instructions in the code, not executed counts)")
    for name, body in zip(names, parts):
        short = re.sub(r"\(.*$", "", name).replace("void ", "")
        if want and not any(w in short for w in want):
            continue
        ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", body, flags=re.M)
        cnt = collections.Counter()
        other = collections.Counter()
        for op in ops:
            if op == "NOP":
                continue
            for cls, pat in CLASSES:
                if re.match(pat, op):
                    cnt[cls] += 1
                    break
            else:
                cnt["other"] += 1
                other[op.split(".")[0]] += 1
        total = sum(cnt.values())
        print(f"\n{short}: {total} instructions")
        for cls, _ in CLASSES + [("other", "")]:
            if cnt[cls]:
                print(f"  {cls:<32} {cnt[cls]:>6}  {100.0 * cnt[cls] / total:5.1f} %")
        if other:
            print("  other = " + ", ".join(f"{k} {v}" for k, v in other.most_common(8)))


if __name__ == "__main__":
    main()
