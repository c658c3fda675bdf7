"""Nsight Compute CSV -> {kernel: {metric: value}} keyed on metric NAMES.

Replaces benchmarks/*/getGpuMetrics.py, which walks a fixed ordered list of 58 metric labels of
Nsight Compute 2020.3 by substring match (getGpuMetrics.py:9) and breaks on any other version.
Works on both `--page raw --csv` (one row per launch, one column per metric) and the default
`--csv` page (one row per launch x metric)."""
from __future__ import annotations

import csv
import io
from typing import Dict, List

# what every tuned configuration is justified with (BASELINE.json north_star)
METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
]

_UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
               "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
               "second": 1.0}


def _num(text: str):
    t = text.replace(",", "").strip()
    try:
        return float(t)
    except ValueError:
        return None


def parse(text: str) -> List[Dict]:
    """One dict per profiled launch: {"kernel": name, "id": n, metric: value-in-base-units, ...}
    (bytes and seconds are normalised; percentages and counts are left as printed)."""
    lines = [l for l in text.splitlines() if l.startswith('"')]
    if not lines:
        return []
    rows = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr = rows[0]
    out: List[Dict] = []
    if "Metric Name" in hdr:                       # long format
        ci = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
        cur = {}
        for r in rows[1:]:
            if len(r) <= ci["Metric Value"]:
                continue
            key = r[ci["ID"]]
            if not cur or cur["id"] != key:
                cur = {"id": key, "kernel": r[ci["Kernel Name"]]}
                out.append(cur)
            v = _num(r[ci["Metric Value"]])
            if v is not None:
                cur[r[ci["Metric Name"]]] = v * _UNIT_SCALE.get(r[ci["Metric Unit"]], 1.0)
    else:                                          # raw page: header row, units row, data rows
        units = rows[1] if len(rows) > 1 else []
        kcol = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        for r in rows[2:]:
            d = {"id": r[0], "kernel": r[kcol] if kcol is not None else ""}
            for i, name in enumerate(hdr):
                if i < len(r) and "__" in name:
                    v = _num(r[i])
                    if v is not None:
                        d[name] = v * _UNIT_SCALE.get(units[i] if i < len(units) else "", 1.0)
            out.append(d)
    return out


def summarise(rows: List[Dict], kernel_substr: str = "dr_") -> Dict:
    """Mean of every metric over the launches of the matching kernel + derived DRAM GB/s."""
    sel = [r for r in rows if kernel_substr in r.get("kernel", "")]
    if not sel:
        return {}
    keys = set().union(*[set(r) for r in sel]) - {"id", "kernel"}
    s = {k: sum(r[k] for r in sel if k in r) / max(1, sum(1 for r in sel if k in r)) for k in keys}
    s["launches"] = len(sel)
    s["kernel"] = sel[0]["kernel"]
    t = s.get("gpu__time_duration.sum")
    if t and "dram__bytes_read.sum" in s and "dram__bytes_write.sum" in s:
        s["dram_bytes"] = s["dram__bytes_read.sum"] + s["dram__bytes_write.sum"]
        s["dram_gbs"] = s["dram_bytes"] / t / 1e9
    return s
