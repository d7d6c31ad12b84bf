#!/usr/bin/env python3
"""Hot CUDA source lines of one captured launch (needs -lineinfo and --import-source on):
   ncu_lines.py report.ncu-rep launch_index [top]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out = []
fname = ""
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" or len(r) < 10: continue   # per-line aggregate rows carry the line number; SASS rows do not
    ix = {}
    for i, h in enumerate(hdr): ix.setdefault(h, i)
    def f(name):
        try:
            return float(r[ix[name]])
        except Exception:
            return 0.0
    out.append((f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed"), fname, r[0], r[1].strip()[:105]))
ts = sum(o[0] for o in out) or 1; ti = sum(o[1] for o in out) or 1
print("total samples %d, warp instructions %.3g" % (ts, ti))
for s, i, t, fn, ln, txt in sorted(out, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  lanes %4.1f  %s:%s  %s" % (100 * s / ts, 100 * i / ti, t / max(i, 1), fn, ln, txt))
