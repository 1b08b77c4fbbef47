#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (tcgen05 MMA / TMEM / TMA / setmaxnreg).
usage: python tools/sass_summary.py [deepof_b200/libdeepof_b200.so] > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys

MNEM = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "USETMAXREG", "ELECT", "SYNCS", "REDG", "REDUX", "BRA.U.ANY", "STL", "LDL"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, counts, sizes = None, collections.OrderedDict(), {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            sizes[cur] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        sizes[cur] += 1
        op = m.group(1)
        for k in MNEM:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# {path}: SASS mnemonic counts per kernel (cuobjdump -sass); kernels with tcgen05 / TMEM / TMA instructions first")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UTMALDG/UBLKCP = TMA tensor / bulk copy,")
    print("# USETMAXREG = setmaxnreg, BRA.U.ANY = waterfall loops around uniform-datapath instructions, STL/LDL = spills")
    rows = []
    for (k, c), name in zip(counts.items(), demangle):
        rows.append((-(c["UTCHMMA"] + c["UTMALDG"] + c["UBLKCP"] + c["LDTM"]), re.sub(r"\(.*", "", name)[:90], sizes[k], c))
    tot = collections.Counter()
    for _, name, n, c in sorted(rows, key=lambda r: (r[0], r[1])):
        tot.update(c)
        if sum(c[m] for m in MNEM[:9]) == 0:
            continue
        print(f"{name:92s} {n:6d} instr  " + " ".join(f"{m}={c[m]}" for m in MNEM if c[m]))
    print("TOTAL " + " ".join(f"{m}={tot[m]}" for m in MNEM if tot[m]) + f"  kernels={len(rows)}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "deepof_b200/libdeepof_b200.so")
