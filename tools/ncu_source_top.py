#!/usr/bin/env python
"""Top stall sites of every captured launch of an `ncu --set full --import-source on` report.
usage: python tools/ncu_source_top.py x.ncu-rep [top_n] > profiles/x_source.txt
Reads `ncu -i x.ncu-rep --page source --csv --print-source sass` and prints, per launch, the SASS instructions with the most
warp-stall samples (with the CUDA source line when -lineinfo resolved it) and the samples summed per source line."""
import collections
import csv
import io
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    # the output is one CSV table per launch, separated by a line with the kernel name
    blocks, cur, name = [], [], None
    for ln in out.splitlines():
        if ln.startswith('"Kernel Name"') or ln.startswith("Kernel Name"):
            if cur:
                blocks.append((name, cur))
            name, cur = ln, []
        else:
            cur.append(ln)
    if cur:
        blocks.append((name, cur))
    for name, lines in blocks:
        rows = list(csv.reader(io.StringIO("\n".join(lines))))
        rows = [r for r in rows if r]
        if not rows:
            continue
        hdr = rows[0]
        print("=" * 100)
        print(name)
        samp = [i for i, h in enumerate(hdr) if "Sampl" in h and "All" in h and "Not" not in h]
        if not samp:
            samp = [i for i, h in enumerate(hdr) if "Sampl" in h]
        src = [i for i, h in enumerate(hdr) if h.strip() in ("Source", "SASS", "Instruction")]
        loc = [i for i, h in enumerate(hdr) if "File" in h or "Line" in h or h.strip() == "Location"]
        print("columns:", hdr[:12], "... samples col:", [hdr[i] for i in samp], "loc:", [hdr[i] for i in loc])
        if not samp or not src:
            continue
        si, ci = samp[0], src[0]
        data = []
        for r in rows[1:]:
            if len(r) <= max(si, ci):
                continue
            try:
                v = float(r[si].replace(",", "") or 0)
            except ValueError:
                continue
            data.append((v, r))
        tot = sum(v for v, _ in data) or 1.0
        per_loc = collections.Counter()
        for v, r in data:
            key = " ".join(r[i] for i in loc) if loc else ""
            per_loc[key] += v
        print(f"total samples {tot:.0f}; top {top} instructions:")
        for v, r in sorted(data, key=lambda t: -t[0])[:top]:
            print(f"  {100 * v / tot:5.1f}%  {r[ci][:90]:90s} {' '.join(r[i] for i in loc)[:60]}")
        if loc:
            print("per source location:")
            for k, v in per_loc.most_common(top):
                print(f"  {100 * v / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
