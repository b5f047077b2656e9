#!/usr/bin/env python
"""Per-kernel summary of ONE training step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py gpurun_out/launches.csv [marker-substring] > profiles/xyz.md
Steps are cut at the first launch of `marker` after a gap of 50 launches (default marker: shard_single_kernel, the
first kernel of the integrated C3 step); the LAST complete step is summarised.  ncu serialises launches and runs them
cold-cache: compare SHARES, not absolute times."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "shard_single_kernel"
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    rows = [(row[idx["Kernel Name"]], float(row[idx["Metric Value"]]), row[idx["Metric Unit"]]) for row in r if len(row) >= len(hdr)]
    names = [n for n, _, _ in rows]
    starts, prev = [], -1000
    for i, n in enumerate(names):
        if marker in n:
            if i - prev > 50:
                starts.append(i)
            prev = i
    if len(starts) >= 2:
        s0, s1 = starts[-2], starts[-1]
        step = rows[s0:s1]
    else:
        # a short capture (`--launch-skip N -c M` with M a little above one step): the steps are identical, so any
        # window of one period holds one step's launches.  The period = the smallest shift that maps the list onto itself.
        period = next((p for p in range(50, len(names) - 4) if names[:len(names) - p] == names[p:]), None)
        if period is None:
            raise SystemExit("no complete step and no period found in the launch list")
        step = rows[:period]
    unit = step[0][2]
    scale = 1e-3 if unit in ("ns", "nsecond") else 1.0
    total = sum(v for _, v, _ in step) * scale
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, _ in step:
        k = re.sub(r"\(.*", "", n)
        k = re.sub(r"^void ", "", k)[:100]
        agg[k][0] += 1
        agg[k][1] += v * scale
    ours = sum(v for k, (c, v) in agg.items() if k.startswith("tt::"))
    print(f"launches in the step: {len(step)}; serialised kernel time {total / 1e3:.3f} ms; tt:: kernels {ours / total * 100:.1f} % of it "
          f"({sum(c for k, (c, v) in agg.items() if k.startswith('tt::'))} launches)\n")
    print("| us | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v:.1f} | {v / total * 100:.1f} % | {c} | `{k}` |")


if __name__ == "__main__":
    main()
