"""Turn ncu exports into the small summaries kept under profiles/.

    python tools/ncu_summary.py raw <raw.csv> <out.json> [note]      # --page raw --csv of a --set full capture -> selected metrics per launch
    python tools/ncu_summary.py launches <launches.csv> <out.csv>    # --metrics gpu__time_duration.sum launch list -> per-kernel shares
"""
import csv
import json
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def raw(src, dst, note=""):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = OrderedDict()
        rec["Kernel Name"] = r[hdr.index("Kernel Name")]
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith(STALLS) and not h.endswith("_not_issued")):
                rec[h] = f"{r[i]} {units[i]}".strip()
        out.append(rec)
    json.dump({"note": note, "kernels": out}, open(dst, "w"), indent=1)
    for rec in out:
        print(json.dumps(rec)[:2000])


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v * 1000.0 if r[ui] in ("ms", "msecond") else v
        a = agg.setdefault(r[ki], [0, 0.0, set()])
        a[0] += 1
        a[1] += v
        a[2].add(f"{r[gi]} x {r[bi]}")
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,share_pct,grids\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{a[0]},{a[1]:.1f},{100 * a[1] / total:.2f},\"{' | '.join(sorted(a[2]))}\"\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"raw": raw, "launches": launches}[sys.argv[1]](*sys.argv[2:])
