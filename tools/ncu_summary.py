#!/usr/bin/env python
"""Markdown summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.md"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("gpc__cycles_elapsed.avg.per_second", "SM clock"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__inst_executed.sum", "warp instructions")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`)\n")
    for r in rows[2:]:
        print(f"## `{r[idx['Kernel Name']][:110]}`\n")
        print("| metric | value |\n|---|---|")
        for k, name in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                print(f"| {name} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        st = [(h, float(r[idx[h]])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and
              h.endswith("_per_issue_active.ratio") and r[idx[h]] not in ("", "n/a")]
        st.sort(key=lambda x: -x[1])
        top = ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                        for h, v in st[:6])
        print(f"\nwarp stall reasons (warps stalled per issue-active cycle): {top}\n")


if __name__ == "__main__":
    main()
