#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of `bench.py --no-graph`:

    python tools/launch_summary.py gpurun_out/launches.csv profiles/rN_launches_last_step.csv profiles/rN_launch_summary.md

The last complete step is the 24 launches that end with the multiclass-NMS emit kernel (the bench's roofline
sub-measurements follow it in the list and are ignored)."""
import csv
import sys


def main():
    src, out_csv, out_md = sys.argv[1:4]
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    rows = [(r[ki], float(r[vi].replace(",", ""))) for r in rd]
    ends = [i for i, (k, _) in enumerate(rows) if "mc_emit_kernel" in k]
    starts = [i for i, (k, _) in enumerate(rows) if "conv_tc_kernel<1" in k]
    end = ends[-1]
    # a step = tower conv ... select_decode, NMS kernels ... emit: walk back to the first launch after the previous emit
    start = ends[-2] + 1 if len(ends) > 1 else starts[0]
    step = rows[start:end + 1]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["order", "kernel", "gpu__time_duration.sum [ns]"])
        for i, (k, v) in enumerate(step):
            w.writerow([i, k, v])
    total = sum(v for _, v in step)
    ours = sum(v for k, v in step if "s2a::" in k)
    agg = {}
    for k, v in step:
        name = k.split("(")[0]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
    with open(out_md, "w") as f:
        f.write("Last complete step = %d launches, %.1f us summed; this library's kernels (`s2a::`) = %.1f us (%.1f %%).\n\n"
                % (len(step), total / 1e3, ours / 1e3, 100.0 * ours / total))
        f.write("| share | us | launches | kernel |\n|---|---|---|---|\n")
        for name, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write("| %.1f %% | %.1f | %d | `%s` |\n" % (100.0 * v / total, v / 1e3, n, name[:90]))
    print("step: %d launches, %.1f us" % (len(step), total / 1e3))


if __name__ == "__main__":
    main()
