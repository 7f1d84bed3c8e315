#!/usr/bin/env python
"""profiles/<name>.txt from the built library: per kernel the SASS instruction count and a histogram of
the mnemonics that matter here (packed FP32, FP64, MUFU, shuffles, TMA bulk copies, async copies), then
the longest straight-line block of the C2 float32 kernel (the step's insolation + balance + statistics).
  python scripts/sass_summary.py enrgy_b200/csrc/libenrgy_b200.so profiles/r02_sass.txt"""
import collections
import re
import subprocess
import sys

WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FMNMX", "FSEL", "FSETP", "MUFU", "DFMA", "DADD", "DMUL", "DSETP",
         "F2F", "SHFL", "VOTE", "LDS", "STS", "LDG", "STG", "UBLKCP", "LDGSTS", "UTMALDG", "SYNCS", "BAR", "BRA", "SHF", "LOP3"]


def main():
    lib, out = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m and name is not None:
            funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
    with open(out, "w") as f:
        f.write("SASS of %s (cuobjdump -sass, sm_100a)\n\n" % lib)
        f.write("%-8s %s\n" % ("instr", "kernel  |  mnemonic histogram"))
        for k, ins in funcs.items():
            if not ins:
                continue
            h = collections.Counter()
            for _, t in ins:
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                mn = t.split()[0].split(".")[0]
                if mn in WATCH:
                    h[mn] += 1
            short = re.sub(r"\(.*", "", k.replace("(anonymous namespace)::", "")).replace("enrgy::", "")
            f.write("%-8d %s\n         %s\n" % (len(ins), short, "  ".join("%s %d" % (m, h[m]) for m in WATCH if h[m])))
        # hot block of the C2 float32 kernel: longest run without a branch / label-target in between
        key = next(k for k in funcs if "energy_balance_kernel<float, 8, 1, false, false, 1, 8, true, 1>" in k)
        ins = funcs[key]
        targets = set()
        for _, t in ins:
            m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
            if m:
                targets.add(int(m.group(1), 16))
        best, cur = [], []
        for a, t in ins:
            if a in targets:
                if len(cur) > len(best):
                    best = cur
                cur = []
            cur.append((a, t))
            if re.match(r"(@!?U?P\d+\s+)?(BRA|EXIT|RET|BSYNC|WARPSYNC)", t):
                if len(cur) > len(best):
                    best = cur
                cur = []
        f.write("\nLongest straight-line block of %s: %d instructions\n" % (re.sub(r"\(.*", "", key), len(best)))
        h = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in best)
        f.write("mnemonics: %s\n\n" % "  ".join("%s %d" % kv for kv in h.most_common()))
        for a, t in best:
            f.write("/*%05x*/  %s\n" % (a, t))
    print("wrote", out)


if __name__ == "__main__":
    main()
