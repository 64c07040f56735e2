#!/usr/bin/env python
"""profiles/r2_k2.md from the round-2 captures: launch-list shares, `--set full` metrics of every K2 kernel, and the
per-opcode stall attribution of the two sweep kernels (ncu source page).

    python profiles/summarize_r2.py profiles/r2_launches.csv profiles/r2_k2_sweeps.ncu-rep gpurun_out/r2_k2_small.ncu-rep"""
import collections
import csv
import re
import subprocess
import sys

launches, rep_sweeps, rep_small = sys.argv[1:4]
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers / thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (realtime) %"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory cycles active %"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM instruction pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
STALLS = ["wait", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "not_selected", "dispatch_stall", "barrier", "mio_throttle",
          "lg_throttle", "branch_resolving", "no_instruction"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")


def launch_shares(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1e-3)
        tot[short(r[ik])] += v
        cnt[short(r[ik])] += 1
    s = sum(tot.values())
    out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append("| `%s` | %d | %.1f | %.1f%% |" % (k, cnt[k], v, 100 * v / s))
    return "\n".join(out)


def raw_table(rep, min_ms=0.2):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(h)}
    dcol = idx["gpu__time_duration.sum"]
    keep = []
    seen = set()
    for r in data:
        ms = float(r[dcol].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(units[dcol], 1e-6)
        n = short(r[idx["Kernel Name"]])
        if ms >= min_ms and n not in seen:
            keep.append(r); seen.add(n)
    names = [short(r[idx["Kernel Name"]])[:30] for r in keep]
    out = ["| metric | " + " | ".join("`%s`" % n for n in names) + " |", "|---|" + "---:|" * len(names)]
    for k, label in KEYS:
        if k in idx:
            out.append("| %s (%s) | " % (label, units[idx[k]]) + " | ".join(r[idx[k]][:12] for r in keep) + " |")
    for st in STALLS:
        k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % st
        if k in idx:
            out.append("| stall %s (cycles / issued instr.) | " % st + " | ".join("%.2f" % float(r[idx[k]]) for r in keep) + " |")
    return "\n".join(out)


def opcode_table(rep, kernel_regex):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex, "--launch-skip", "0",
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    ia, isrc, iss, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    seen, per = set(), collections.defaultdict(lambda: [0, 0])
    tot_ex = tot_s = 0
    for r in rows[2:]:
        if len(r) <= iex or r[ia] in seen:
            continue
        seen.add(r[ia])
        try:
            s, ex = int(r[iss]), int(r[iex])
        except ValueError:
            continue
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        per[op][0] += ex; per[op][1] += s
        tot_ex += ex; tot_s += s
    out = ["| opcode | executed (% of warp instructions) | stall samples % |", "|---|---:|---:|"]
    for op, (ex, s) in sorted(per.items(), key=lambda x: -x[1][0])[:16]:
        out.append("| `%s` | %.1f | %.1f |" % (op, 100.0 * ex / tot_ex, 100.0 * s / max(tot_s, 1)))
    return "\n".join(out)


git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
md = ["# ncu summary — round 2, final build (captured at git %s)" % git, "",
      "Per-launch times below are cold-cache and serialised (ncu replay): compare SHARES, not absolutes.  The numbers quoted in",
      "DESIGN.md and printed by bench.py come from CUDA events, never from a run under the profiler.", "",
      "## Launch list shares (`%s`: `bench.py --steps 2 --warmup 3`, launches 60..460)" % launches, "", launch_shares(launches), "",
      "## `--set full` capture of the sweep kernels (`%s`; `tools/k2_time.py --pairs 2048`: one launch = 2048 C3-shaped pairs, 8.15e9 cells)" % rep_sweeps,
      "", raw_table(rep_sweeps), "",
      "Per-pair constants derived from this report: `profiles/k2_constants.json` (`profiles/make_constants.py`).", "",
      "### Where the emit sweep's instructions and stalls go (source page, per opcode)", "", opcode_table(rep_sweeps, "tc_emit_kernel|fast_emit_kernel"), "",
      "### The same for the column histogram sweep", "", opcode_table(rep_sweeps, "tc_hist_kernel|fast_hist_kernel"), "",
      "## `--set full` capture of the other K2 kernels and K3 (same command)", "", raw_table(rep_small, 0.05), ""]
open("profiles/r2_k2.md", "w").write("\n".join(md))
print("wrote profiles/r2_k2.md")
