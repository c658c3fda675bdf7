"""Turns the Nsight Compute output a gpurun call brought back (gpurun_out/) into the committed
evidence under profiles/: per-kernel summaries (time, DRAM bytes, throughput, pipe utilisation,
occupancy limits), the launch list of the bench command with each kernel's share, and
profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).

    python tools/make_profiles.py <round tag, e.g. r01>
"""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drstencil_b200.tuner import metrics  # noqa: E402

KEEP = metrics.METRICS + [
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_issued.sum", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "l1tex__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    traffic_path = os.path.join(out_dir, "traffic.json")
    traffic = {}          # workload -> DRAM bytes per launch, rebuilt from this run's reports
    summary = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "*.ncu-rep"))):
        name = os.path.splitext(os.path.basename(rep))[0]
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        rows = metrics.parse(r.stdout)
        if not rows:
            continue
        s = metrics.summarise(rows)
        if not s:
            continue
        keep = {k: s[k] for k in KEEP + ["kernel", "launches", "dram_bytes", "dram_gbs"] if k in s}
        summary[name] = keep
        # keyed by workload (prof_<workload>.ncu-rep): c4 and c4t2 share a kernel NAME but not a kernel
        traffic[name.split("_", 1)[1] if "_" in name else name] = s.get("dram_bytes")
    for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "dram_*.csv"))):
        rows = metrics.parse(open(f).read())
        s = metrics.summarise(rows)
        if s:
            summary[os.path.splitext(os.path.basename(f))[0]] = {k: v for k, v in s.items() if not isinstance(v, dict)}
            traffic[os.path.splitext(os.path.basename(f))[0].split("_", 1)[1]] = s.get("dram_bytes")
    json.dump(summary, open(os.path.join(out_dir, "%s_ncu_summary.json" % tag), "w"), indent=1, sort_keys=True)
    # the same as a table
    with open(os.path.join(out_dir, "%s_ncu_summary.md" % tag), "w") as f:
        f.write("# %s -- Nsight Compute summary of the sweep kernels (B200, `ncu --set full --clock-control none --import-source on "
                "-k regex:dr_ -s 3 -c 2` on `python -m drstencil_b200.tuner.run_one ...`, each only after the same command had "
                "exited 0 without ncu; mean of the profiled launches)\n\n" % tag)
        f.write("`prof_<preset>`: the preset's kernel on its own grid; `prof_c5slab`: the c5 preset on a 256 x 1536 x 1536 slab (54 GiB of "
                "state is too much for full-set replays); `dram_c5`: DRAM bytes and duration of the c5 launches inside the bench command.\n\n")
        f.write("| report | kernel | time | DRAM read | DRAM written | DRAM GB/s | DRAM % of peak | L2 bytes | L2 % | shared wavefronts | bank conflicts | "
                "FP64 pipe | FMA pipe | issue active | warps active | regs | occupancy limit regs / smem (CTAs) |\n|" + "---|" * 17 + "\n")
        g = lambda d, k, scale=1.0, fmt="%.1f": (fmt % (d[k] * scale)) if k in d else "-"
        for name, d in sorted(summary.items()):
            f.write("| `%s` | `%s` | %s ms | %s GB | %s GB | %s | %s %% | %s GB | %s %% | %s | %s | %s %% | %s %% | %s %% | %s %% | %s | %s / %s |\n" % (
                name, d.get("kernel", "")[:40], g(d, "gpu__time_duration.sum", 1e3, "%.4f"), g(d, "dram__bytes_read.sum", 1e-9, "%.3f"),
                g(d, "dram__bytes_write.sum", 1e-9, "%.3f"), g(d, "dram_gbs", 1.0, "%.0f"),
                g(d, "dram__throughput.avg.pct_of_peak_sustained_elapsed"), g(d, "lts__t_bytes.sum", 1e-9, "%.2f"),
                g(d, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), g(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1.0, "%.3g"),
                g(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0, "%.3g"),
                g(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), g(d, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                g(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"), g(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                g(d, "launch__registers_per_thread", 1.0, "%.0f"), g(d, "launch__occupancy_limit_registers", 1.0, "%.0f"),
                g(d, "launch__occupancy_limit_shared_mem", 1.0, "%.0f")))
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    # launch list of the bench command: per-kernel totals and shares
    ll = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if os.path.exists(ll):
        rows = metrics.parse(open(ll).read())
        agg = {}
        for r in rows:
            k = r["kernel"][:80]
            t = r.get("gpu__time_duration.sum", 0.0)
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += t
        total = sum(a[1] for a in agg.values())
        with open(os.path.join(out_dir, "%s_bench_launches.md" % tag), "w") as f:
            f.write("# ncu launch list of `python bench.py --no-extras --no-parity --steps 1 --warmup 3` (first %d launches)\n\n" % len(rows))
            f.write("cold-cache, serialised times: compare SHARES, not absolutes\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
            for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, n, t * 1e3, 100 * t / total))
        import shutil
        shutil.copy(ll, os.path.join(out_dir, "%s_bench_launches.csv" % tag))
    print(json.dumps(summary, indent=1)[:3000])


if __name__ == "__main__":
    main()
