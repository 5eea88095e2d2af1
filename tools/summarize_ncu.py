#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and launch-list CSVs into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py report  gpurun_out/prof_fit_r1.ncu-rep  profiles/r1_fit_kernel.md  "title"
    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv     profiles/r1_launches.md    "title"
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def report(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `%s` (`ncu --set full --clock-control none --import-source on`), one column per captured launch.\n\n" % (title, rep))
        launches = rows[2:]
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(launches))) + " |\n|---|---|" + "---|" * len(launches) + "\n")
        f.write("| kernel | | " + " | ".join(re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("void hpsdf::", "") for r in launches) + " |\n")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                f.write("| `%s` | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in launches)))
        # SASS opcode mix of the first launch
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        hi = [i for i, r in enumerate(srows) if "Source" in r and "Instructions Executed" in r]
        if hi:
            hh = srows[hi[0]]
            si, ei = hh.index("Source"), hh.index("Instructions Executed")
            agg = collections.Counter()
            end = hi[1] if len(hi) > 1 else len(srows)
            for r in srows[hi[0] + 1:end]:
                try:
                    n = int(float(r[ei]))
                except (ValueError, IndexError):
                    continue
                m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
                if m:
                    agg[m.group(2)] += n
            tot = sum(agg.values())
            f.write("\n## Executed warp instructions by SASS opcode (launch 0)\n\n| opcode | warp instructions | share |\n|---|---|---|\n")
            for k, v in agg.most_common(16):
                f.write("| %s | %d | %.1f %% |\n" % (k, v, 100.0 * v / tot))
            tensor = [k for k in agg if k.startswith(("UTC", "HMMA", "DMMA", "LDTM", "STTM", "UTMA"))]
            f.write("\nTensor / TMA opcodes present: %s (FP64 work on CUDA cores; tcgen05 has no FP64 kind).\n" % (tensor or "none"))


def launches(csvfile, out, title):
    rows = list(csv.reader(open(csvfile, errors="ignore")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    h = rows[hdr[0]]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr[0] + 2:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("hpsdf::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `%s` (`ncu --metrics gpu__time_duration.sum --clock-control none`): per-launch times are cold-cache and "
                "serialised, so only the SHARES are meaningful.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % (title, csvfile))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k, v[0], v[1] / 1e3, 100.0 * v[1] / tot))


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
