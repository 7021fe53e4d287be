#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table:
per (kernel, grid) the launch count, mean duration and the share of all captured device time.

    python scripts/summarize_launches.py gpurun_out/launches_r1c.csv > profiles/r1c_launches.md

ncu serialises launches and runs them cold-cache, so the SHARES are what to compare with
bench.py's CUDA-event numbers, not the absolute times (B200_PROFILING.md)."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*$", "", name)            # drop the argument list
    name = name.replace("void ", "").replace("pulpo::", "")
    return name.strip()


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    k_i, g_i, b_i, v_i, u_i = (hdr.index(c) for c in ("Kernel Name", "Grid Size", "Block Size", "Metric Value", "Metric Unit"))
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= v_i:
            continue
        ns = float(r[v_i].replace(",", ""))
        if r[u_i] in ("us", "usecond"):
            ns *= 1e3
        key = (short(r[k_i]), r[g_i], r[b_i])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    ours = sum(v[1] for k, v in agg.items() if not k[0].startswith("at::"))
    print("# ncu launch list summary: %s" % path)
    print()
    print("%d launches captured, %.1f us total device time, %.1f us (%.1f %%) in libpulpo_b200 kernels."
          % (sum(v[0] for v in agg.values()), total / 1e3, ours / 1e3, 100 * ours / max(total, 1)))
    print("Times are per-launch under ncu (cold cache, serialised): compare shares, not absolutes.")
    print()
    print("| kernel | grid | block | launches | mean us | total us | share |")
    print("|---|---|---|---:|---:|---:|---:|")
    for (name, grid, block), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f %% |" % (name, grid, block, n, ns / n / 1e3, ns / 1e3, 100 * ns / total))


if __name__ == "__main__":
    main(sys.argv[1])
