#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches_iqap_b1024.md [first last]
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_full_<name>.md

`launches`: per-kernel totals / averages / shares of one bench step from the
`--metrics gpu__time_duration.sum` launch list (cold-cache, serialised: compare SHARES with bench.py's live
profile, not absolutes).  `full`: the metrics the roofline needs from an `ncu --set full` capture (DRAM bytes,
DRAM / tensor-pipe utilisation, occupancy, registers) and writes dram traffic per launch into profiles/traffic.json.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CLASS_OF = [  # kernel-name fragment -> bench.py kernel class (best effort; GEMM classes are told apart by template args)
    ("row_attn_kernel", "dec_self_attention / dec_cross_attention"),
    ("enc_attention_kernel", "enc_attention"),
    ("ffn_partial_kernel", "dec_ffn_split"),
    ("ffn_reduce_ln_kernel", "dec_ffn_split"),
    ("gemm_tc_kernel", "gemm family"),
]


def short(name):
    name = name.split("(")[0]
    return name.replace("b200vqa::", "").replace("<unnamed>::", "").replace("void ", "").replace("unnamed>::", "")[-64:]


def launches(path, out, first=None, last=None):
    """first / last: launch index range [first, last) of ONE bench step inside the list (steps start at the
    embedding kernel); default = the whole list."""
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    body = [r for r in rows[start + 1:] if len(r) > vi]
    if first is not None:
        body = body[first:last]
    for r in body:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({os.path.basename(path)})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over the launches of timed bench steps; "
                "times are cold-cache and serialised (no PDL / branch overlap), so compare shares, not absolutes.\n\n")
        f.write(f"launches: {sum(a[0] for a in agg.values())}, total {total / 1e3:.3f} ms\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.2f} | {a[1] / total:.3f} |\n")
    print(open(out).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__cycles_elapsed.max"]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(val.replace(",", "")) * mult


def full(path, out, tag=None):
    """path: an .ncu-rep, or the `ncu -i rep --page raw --csv` text already exported on the GPU box (reports with many
    launches exceed what gpurun copies back)."""
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True,
                             check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(path)})\n\n")
        seen = set()
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            key = (name, r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "")
            if key in seen:
                continue
            seen.add(key)
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            rd = wr = None
            for m in WANT:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
                    if m == "dram__bytes_read.sum":
                        rd = to_bytes(r[i], units[i])
                    if m == "dram__bytes_write.sum":
                        wr = to_bytes(r[i], units[i])
            if rd is not None and wr is not None:
                f.write(f"| dram traffic per launch | {rd + wr:.0f} | byte |\n")
                traffic.setdefault(name, []).append(rd + wr)
            f.write("\n")
    print(open(out).read())
    if tag:
        tpath = os.path.join(REPO, "profiles", "traffic.json")
        cur = json.load(open(tpath)) if os.path.exists(tpath) else {}
        vals = [v for vs in traffic.values() for v in vs]
        cur[tag] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_sampled": len(vals),
                    "source": os.path.basename(out)}
        json.dump(cur, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], *(int(a) for a in sys.argv[4:6]))
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
