"""Rank the source lines of a kernel by warp-stall samples: reads `ncu -i REP --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_hot_lines.py REP.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
samples, insts, text = collections.Counter(), collections.Counter(), {}
fname, H = "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        H = r
        si, ii = H.index("Warp Stall Sampling (All Samples)"), H.index("Instructions Executed")
    elif H and len(r) > max(si, ii) and r[0].isdigit():
        key = (fname, int(r[0]))
        if r[2] == "-":          # the source line itself (SASS rows carry the address)
            text[key] = r[1].strip()[:120]
        try:
            if r[2] == "-":
                samples[key] += int(r[si]); insts[key] += int(r[ii])
        except ValueError:
            pass
tot = sum(samples.values())
print("total samples", tot)
for k, v in samples.most_common(top):
    print(f"{100 * v / tot:5.1f}%  warp-insts {insts[k]:10d}  {k[0]}:{k[1]:<5d} {text.get(k, '')}")
