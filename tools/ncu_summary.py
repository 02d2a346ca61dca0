#!/usr/bin/env python
"""Turn ncu artefacts from gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md>
        per-kernel launch list (count, mean/min/max gpu__time_duration, share of the total)
    python tools/ncu_summary.py report <file.ncu-rep> <out.md> [--json profiles/ncu_summary.json]
        key counters of every kernel in one `ncu --set full` capture; --json merges
        {kernel: {dram_bytes_per_launch, ...}} into the file bench.py reads `roofline.traffic` from

Needs the `ncu` binary (to read .ncu-rep); runs without a GPU.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
]

_UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short_name(k):
    k = re.sub(r"^void ", "", k)
    return re.sub(r"\(.*$", "", k)


def launches(csv_path, out_md):
    lines = [l for l in open(csv_path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1e-3)
        per.setdefault((short_name(r["Kernel Name"]), r["Block Size"], r["Grid Size"]), []).append(v * scale)
    total = sum(sum(v) for v in per.values())
    with open(out_md, "w") as f:
        f.write("# ncu launch list — `%s`\n\n" % os.path.basename(csv_path))
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised "
                "launches: shares are meaningful, absolutes are not bench numbers).\n\n")
        f.write("| kernel | block | grid | launches | mean us | min us | max us | share of listed time |\n")
        f.write("|---|---|---|---|---|---|---|---|\n")
        for (k, b, g), v in per.items():
            f.write("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f | %.1f %% |\n" % (
                k, b, g, len(v), sum(v) / len(v), min(v), max(v), 100.0 * sum(v) / total))
        f.write("\nTotal listed GPU time: %.1f us over %d launches.\n" % (total, sum(len(v) for v in per.values())))


FP64_OPS = ("DFMA", "DADD", "DMUL", "DSETP")


def source_counts(rep_path):
    """Executed warp instructions per warp of the (single) kernel in a capture taken with --import-source on:
    FP64-pipe arithmetic (DFMA / DADD / DMUL / DSETP: each holds the dispatch port for two cycles) and the rest."""
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next((r for r in rows if "Instructions Executed" in r and "Source" in r), None)
    if hdr is None:
        return None
    i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
    first, f64, oth = None, 0, 0
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= max(i_src, i_ex) or not r[i_ex].isdigit():
            continue
        toks = r[i_src].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        n = int(r[i_ex])
        if first is None:
            first = n          # the kernel's first instruction: executed once by every warp
        if op in FP64_OPS:
            f64 += n
        else:
            oth += n
    if not first:
        return None
    return {"warps": first, "fp64_inst_per_warp": f64 / first, "other_inst_per_warp": oth / first,
            "issue_cycles_per_warp": (2 * f64 + oth) / first, "source_page": os.path.basename(rep_path)}


def report(rep_path, out_md, json_path=None):
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    summary = {}
    with open(out_md, "w") as f:
        f.write("# ncu `--set full --clock-control none` — `%s`\n\n" % os.path.basename(rep_path))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = short_name(d["Kernel Name"])
            f.write("## `%s`  (block %s, grid %s)\n\n| metric | value | unit |\n|---|---|---|\n" % (
                name, d.get("Block Size", "?"), d.get("Grid Size", "?")))
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write("| %s | %s | %s |\n" % (k, d[k], u[k]))
            f.write("\n")
            try:
                rd = float(d["dram__bytes_read.sum"].replace(",", "")) * _UNIT_SCALE[u["dram__bytes_read.sum"]]
                wr = float(d["dram__bytes_write.sum"].replace(",", "")) * _UNIT_SCALE[u["dram__bytes_write.sum"]]
                base = re.sub(r"<.*$", "", name)
                summary[base] = {
                    "kernel": name, "dram_bytes_per_launch": rd + wr,
                    "dram_bytes_read": rd, "dram_bytes_write": wr,
                    "fp64_pipe_pct_elapsed": float(d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan")),
                    "issue_active_pct": float(d.get("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "nan")),
                    "duration": d.get("gpu__time_duration.sum") + " " + u.get("gpu__time_duration.sum", ""),
                    "source": os.path.basename(rep_path)}
            except (KeyError, ValueError):
                pass
    if len(summary) == 1:
        sc = source_counts(rep_path)
        if sc:
            for v in summary.values():
                v.update(sc)
            with open(out_md, "a") as f:
                f.write("## executed instructions per warp (source page)\n\n| | |\n|---|---|\n")
                for k in ("warps", "fp64_inst_per_warp", "other_inst_per_warp", "issue_cycles_per_warp"):
                    f.write("| %s | %.1f |\n" % (k, sc[k]))
                f.write("\nissue_cycles_per_warp = 2 x FP64 + other (an FP64 instruction holds the sub-partition's dispatch "
                        "port for two cycles); x warps / 592 sub-partitions = the kernel's elapsed cycles to ~1 %.\n")
    if json_path:
        old = {}
        if os.path.exists(json_path):
            with open(json_path) as jf:
                old = json.load(jf)
        old.update(summary)
        with open(json_path, "w") as jf:
            json.dump(old, jf, indent=1, sort_keys=True)
    return summary


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif len(sys.argv) >= 4 and sys.argv[1] == "report":
        jp = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
        print(json.dumps(report(sys.argv[2], sys.argv[3], jp), indent=1))
    else:
        sys.exit(__doc__)
