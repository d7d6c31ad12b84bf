#!/usr/bin/env python3
"""Summarise an .ncu-rep: per captured launch the headline metrics; optional per-SASS-region breakdown.
usage: ncu_summary.py report.ncu-rep [launch_index_for_sass]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
keys = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc/SM"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("lts__t_bytes.sum", "L2_bytes"), ("lts__t_bytes.sum.per_second", "L2_B/s"),
        ("dram__bytes_read.sum.per_second", "dram_rd/s"), ("l1tex__t_bytes.sum", "L1_bytes"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe%"), ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%")]
units = rows[1]
for n, r in enumerate(rows[2:]):
    print("[%d] %s" % (n, r[hdr.index("Kernel Name")][:70]))
    print("    " + "  ".join("%s=%s%s" % (lab, r[hdr.index(k)][:10], units[hdr.index(k)] if lab in ("time", "dram_rd", "dram_wr", "L2_bytes", "L2_B/s", "dram_rd/s", "L1_bytes") else "") for k, lab in keys if k in hdr))
    st = [(float(r[i] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
    tot = sum(v for v, _ in st) or 1
    print("    stalls: " + "  ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:7]))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [n for n, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[hi[0]]; ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else None)] if len(r) >= len(hdr) - 2]
    f = lambda x: float(x) if x not in ("", None) else 0.0
    tot = sum(f(r[ix["Instructions Executed"]]) for r in data) or 1
    ts = sum(f(r[ix["# Samples"]]) for r in data) or 1
    print("SASS: %d instructions, %.3g warp-inst executed" % (len(data), tot))
    for b in range(0, len(data), 100):
        blk = data[b:b + 100]
        e = sum(f(r[ix["Instructions Executed"]]) for r in blk); t = sum(f(r[ix["Thread Instructions Executed"]]) for r in blk); s = sum(f(r[ix["# Samples"]]) for r in blk)
        if e / tot > 0.01 or s / ts > 0.01:
            print("  sass %5d-%5d  exec %5.1f%%  samples %5.1f%%  threads/inst %.1f" % (b, b + len(blk), 100 * e / tot, 100 * s / ts, t / max(e, 1)))
