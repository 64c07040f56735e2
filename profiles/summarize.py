#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into the tracked files under profiles/.

    python profiles/summarize.py <tag> <launches.csv> <report.ncu-rep> [bench.json]

launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
report       : ncu --set full --clock-control none --import-source on -o ...
"""
import csv
import json
import subprocess
import sys


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    tot, cnt = {}, {}
    for r in rows[hdr + 1:]:
        if len(r) < 10:
            continue
        name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        t = float(r[-1].replace(",", ""))
        tot[name] = tot.get(name, 0.0) + t
        cnt[name] = cnt.get(name, 0) + 1
    s = sum(tot.values())
    out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append("| `%s` | %d | %.1f | %.1f%% |" % (k, cnt[k], v / 1e3, 100 * v / s))
    return "\n".join(out)


KEYS = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs/thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected")]


def raw_table(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(h)}
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:34] for r in data]
    out = ["| metric | " + " | ".join("`%s`" % n for n in names) + " |", "|---|" + "---:|" * len(names)]
    for k, label in KEYS:
        if k not in idx:
            continue
        u = units[idx[k]]
        out.append("| %s (%s) | " % (label, u) + " | ".join(r[idx[k]] for r in data) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    tag, launches, rep = sys.argv[1:4]
    md = ["# ncu summary — %s" % tag, "",
          "Per-launch times below are cold-cache and serialised (ncu replay): compare SHARES, not absolutes.", "",
          "## Launch list shares (`%s`)" % launches, "", launch_shares(launches), "",
          "## `--set full` capture of the hot kernels (`%s`)" % rep, "", raw_table(rep), ""]
    if len(sys.argv) > 4:
        d = [json.loads(l) for l in open(sys.argv[4]) if l.strip().startswith("{")][-1]
        md += ["## bench.py line of the same build", "", "```json", json.dumps(d, indent=1), "```", ""]
    open("profiles/%s.md" % tag, "w").write("\n".join(md))
    print("wrote profiles/%s.md" % tag)
