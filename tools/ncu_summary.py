#!/usr/bin/env python3
"""Summarise an .ncu-rep: per captured launch the headline metrics; optional per-SASS-region breakdown.
usage: ncu_summary.py report.ncu-rep [launch_index_for_sass] [--json profiles/ncu_table.json --key "<scene>:<mode>" --source <file kept under profiles/>]

With --json the heaviest captured launch of every kernel class is merged into the table bench.py reads for its roofline
object (dram bytes per launch = `traffic`, issue-slot utilisation, lanes per instruction, cache hit rates)."""
import csv, subprocess, sys, io, json, os
argv = sys.argv[1:]
opt = {}
for flag in ("--json", "--key", "--source"):
    if flag in argv:
        i = argv.index(flag)
        opt[flag] = argv[i + 1]
        del argv[i:i + 2]
sys.argv = [sys.argv[0]] + argv
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


def _bytes(val, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val or 0) * scale.get(unit.split("/")[0], 1.0)


def _klass(name):
    if "k_shadow_wave" in name: return "k_shadow_wave"
    if "k_shade" in name: return "k_shade"
    if "k_extend" in name: return "k_extend<primary>" if ("<1" in name or "<true" in name or "(bool)1" in name) else "k_extend<queue>"
    return name.split("(")[0].replace("void ", "")


if "--json" in opt:
    table = {}
    if os.path.exists(opt["--json"]):
        table = json.load(open(opt["--json"]))
    entry = table.setdefault(opt.get("--key", "unknown"), {})
    col = lambda r, k: r[hdr.index(k)] if k in hdr else ""
    best = {}
    for r in rows[2:]:
        kl = _klass(r[hdr.index("Kernel Name")])
        t = float(col(r, "gpu__time_duration.sum") or 0)
        if kl not in best or t > best[kl][0]:
            best[kl] = (t, r)
    for kl, (t, r) in best.items():
        tu = units[hdr.index("gpu__time_duration.sum")]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
        rd = _bytes(col(r, "dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")])
        wr = _bytes(col(r, "dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")])
        entry[kl] = {"dram_bytes_per_launch": rd + wr, "launch_seconds_under_ncu": t_s, "dram_gbs": (rd + wr) / t_s / 1e9 if t_s else None,
                     "issue_pct": float(col(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed") or col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0),
                     "threads_per_inst": float(col(r, "smsp__thread_inst_executed_per_inst_executed.ratio") or 0),
                     "ipc_per_sm": float(col(r, "sm__inst_executed.avg.per_cycle_elapsed") or 0),
                     "l1_hit_pct": float(col(r, "l1tex__t_sector_hit_rate.pct") or 0), "l2_hit_pct": float(col(r, "lts__t_sector_hit_rate.pct") or 0),
                     "occupancy_pct": float(col(r, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0),
                     "registers": int(float(col(r, "launch__registers_per_thread") or 0)),
                     "source": opt.get("--source", os.path.basename(rep))}
    json.dump(table, open(opt["--json"], "w"), indent=1, sort_keys=True)

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
