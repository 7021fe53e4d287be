#!/usr/bin/env python
"""Summarise `ncu --set full` reports (read here with `ncu -i ... --page raw --csv`) into a markdown
table plus a JSON of per-launch DRAM traffic that bench.py reports as `roofline.traffic`.

    python scripts/summarize_ncu_full.py TAG rep1.ncu-rep [rep2.ncu-rep ...]
      -> profiles/TAG_ncu_full.md, profiles/TAG_traffic.json
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def short(name):
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("pulpo::", "")
    return name.strip()


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    lines = ["# ncu --set full summaries (%s)" % tag, "",
             "One row per captured launch (`--clock-control none`, cold cache, kernel timed alone). `top stalls` = warp",
             "stall reasons per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active).", ""]
    hdr = ["report", "kernel", "grid x block"] + [k[1] for k in KEYS] + ["top stalls"]
    lines.append("| " + " | ".join(hdr) + " |")
    lines.append("|" + "---|" * len(hdr))
    traffic = {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, units = rows[0], rows[1]
        stall = [c for c in h if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("per_issue_active.ratio")]
        for r in rows[2:]:
            name = short(r[h.index("Kernel Name")])
            cells = [os.path.basename(rep), "`%s`" % name, "%s x %s" % (r[h.index("launch__grid_size")], r[h.index("launch__block_size")])]
            for k, _ in KEYS:
                if k in h:
                    i = h.index(k)
                    v = r[i]
                    try:
                        v = "%.4g" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                    cells.append("%s %s" % (v, units[i]) if units[i] and units[i] != "%" else v)
                else:
                    cells.append("-")
            st = sorted(((float(r[h.index(s)]), s.split("stalled_")[1].split("_per_")[0]) for s in stall), reverse=True)[:3]
            cells.append(", ".join("%s %.1f" % (n, v) for v, n in st))
            lines.append("| " + " | ".join(cells) + " |")
            ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            t = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            key = name
            traffic.setdefault(key, []).append(t)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", "%s_ncu_full.md" % tag), "w").write("\n".join(lines) + "\n")
    json.dump({k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)} for k, v in traffic.items()},
              open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
