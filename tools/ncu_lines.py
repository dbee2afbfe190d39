#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples")
        try:
            lines.append((int(r[i_inst]), int(r[i_s]), fname, r[0], r[1]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines); tots = sum(l[1] for l in lines)
print(f"total warp-instructions {tot}  samples {tots}")
for n, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{n:>11d} {100*n/max(tot,1):5.1f}%  samp {100*s/max(tots,1):5.1f}%  {f}:{ln:>4s} | {src[:100]}")
