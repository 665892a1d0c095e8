"""Summarise ncu outputs under gpurun_out/ into tracked files under profiles/.

    python tools/profile_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md "<command>"
    python tools/profile_summary.py kernel  gpurun_out/prof.ncu-rep profiles/r01_kernel.md "<note>"
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg[r[ki].split("(")[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\ncommand: `{cmd}`\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | mean us | total ms | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f}% |\n")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def kernel(src, dst, note):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary\n\nsource: `{src}`\n\n{note}\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name.split('(')[0]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            stalls = [(h, float(r[i] or 0)) for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
            if stalls:
                f.write("\nwarp stall reasons (average warps stalled per issue-active cycle):\n\n")
                for h, v in sorted(stalls, key=lambda x: -x[1])[:8]:
                    f.write(f"* {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
