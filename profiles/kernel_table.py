#!/usr/bin/env python
"""Per-kernel table (markdown) from an `ncu --page raw --csv` dump of tools/all_kernels.py:
one row per distinct kernel (launches of the same kernel are summed: time, DRAM bytes; rates averaged
weighted by time).  Columns: launches, total time, DRAM read+write, DRAM GB/s and its fraction of the
measured copy peak (MEASURED_PEAKS.json hbm_gbs, fallback 6553.9), L1 and L2 sector hit rates,
shared-memory wavefronts, registers, achieved occupancy, issue-slot utilisation, fp64 pipe."""
import csv, json, re, sys
from collections import OrderedDict
from pathlib import Path

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
try:
    PEAK = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    PEAK = 6553.9


def val(r, name, scale_unit=True):
    if name not in col or r[col[name]] in ("", "n/a"):
        return 0.0
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    if scale_unit:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
              "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return v


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("cub::CUB_300001_SM_1000::", "cub::").replace("cub::detail::scan::", "cub::")[:70]


agg = OrderedDict()
for r in rows[2:]:
    k = short(r[col["Kernel Name"]])
    t = val(r, "gpu__time_duration.sum")
    a = agg.setdefault(k, dict(n=0, t=0.0, rd=0.0, wr=0.0, l1=0.0, l2=0.0, smw=0.0, regs=0, occ=0.0, issue=0.0, fp64=0.0))
    a["n"] += 1; a["t"] += t
    a["rd"] += val(r, "dram__bytes_read.sum"); a["wr"] += val(r, "dram__bytes_write.sum")
    a["l1"] += t * val(r, "l1tex__t_sector_hit_rate.pct"); a["l2"] += t * val(r, "lts__t_sector_hit_rate.pct")
    a["smw"] += val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    a["regs"] = int(val(r, "launch__registers_per_thread"))
    a["occ"] += t * val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["issue"] += t * val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["fp64"] += t * val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
print(f"| kernel | launches | time ms | DRAM GB (rd+wr) | DRAM GB/s | % of {PEAK:.0f} GB/s | L1 hit % | L2 hit % | smem wavefronts | regs | occupancy % | issue % | fp64 pipe % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, a in agg.items():
    t = a["t"] or 1e-30
    gb = (a["rd"] + a["wr"]) / 1e9
    gbs = gb / t
    print(f"| `{k}` | {a['n']} | {a['t'] * 1e3:.3f} | {gb:.3f} | {gbs:.0f} | {100 * gbs / PEAK:.1f} | {a['l1'] / t:.1f} | {a['l2'] / t:.1f} | "
          f"{a['smw']:.3g} | {a['regs']} | {a['occ'] / t:.1f} | {a['issue'] / t:.1f} | {a['fp64'] / t:.1f} |")
