#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table
(markdown).  The last full training step is delimited by the Adam kernel (two per step)."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^(]*>)?)", name)
    return (m.group(1) if m else name)[:90]


def main(path, out):
    rows = []
    with open(path) as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        hdr = next(rd)
        ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
        for r in rd:
            rows.append((r[ki], float(r[vi].replace(",", "")), r[gi], r[bi]))
    adam = [i for i, r in enumerate(rows) if "adam_kernel" in r[0] or "adam_dev_kernel" in r[0]]
    if len(adam) >= 4:
        lo, hi = adam[-3] + 1, adam[-1] + 1          # last full step: after the previous step's G-phase Adam
    else:
        lo, hi = 0, len(rows)
    step = rows[lo:hi]
    agg = OrderedDict()
    for name, ns, grid, block in step:
        a = agg.setdefault(short(name), [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if not k.startswith("at::") and "nccl" not in k.lower())
    with open(out, "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % path)
        f.write("Launches %d..%d of %d = one full WGAN-GP step (B=64, N=2048).  Times are ncu per-launch device\n"
                "durations (cold-cache, serialised): compare SHARES, not absolutes.\n\n" % (lo, hi, len(rows)))
        f.write("total %.3f ms over %d launches; kernels of libspgan_b200: %.1f %% of device time, torch-internal "
                "(autograd accumulate / views): %.1f %%\n\n" % (total / 1e6, len(step), 100 * ours / total, 100 * (1 - ours / total)))
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, a[0], a[1] / 1e6, 100 * a[1] / total))
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
