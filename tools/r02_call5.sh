#!/bin/bash
# Round 2, GPU call 5 (one GPU): TMA start-alignment micro-probe, the full GPU test suite, the contract bench (both arms).
O=gpurun_out/r02_call5
mkdir -p $O
for a in "8 0" "8 2" "8 1" "8 3" "8 -2" "8 -1" "4 4" "4 1" "4 2" "4 -1" "8 0 1" "8 0 1 s" "8 1 1 s"; do
  timeout 60 tools/bin/tma_align_probe $a >> $O/tma_align.txt 2>&1 || echo "  (rc=$? for: $a)" >> $O/tma_align.txt
done
cat $O/tma_align.txt
t0=$(date +%s)
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? seconds=$(( $(date +%s) - t0 ))" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
t1=$(date +%s)
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$? seconds=$(( $(date +%s) - t1 ))"
t2=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$? seconds=$(( $(date +%s) - t2 ))"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_call5/bench.json"))
print("c5 value %.1f frac %.3f ms/step %.1f launches %d clocks %s finite %s" % (d["value"], d["roofline"]["frac"], d["ms_per_step"], d["gpu_launches"], d["clocks"], d.get("data_finite")))
print("e2e %.1f ms/step %.1f bound %s" % (d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"]["bound"].items()}))
print("parity", json.dumps(d["parity"])[:700])
print("cpu", d["cpu_baseline"]["cores"], round(d["cpu_baseline"]["value"], 2))
for e in d["per_config"]:
    if "error" in e:
        print("  ", e); continue
    print("   %-8s %.1f GStencil/s frac %.3f  %d steps %.2f s  clocks %s finite %s cpu %s ref %s" % (e["workload"][:8], e["value"], e["roofline"]["frac"], e["steps"], e["timed_seconds"],
          (e["clocks"]["sm_mhz"], e["clocks"]["reasons"]), e["data_finite"], round(e.get("cpu_baseline", {}).get("value", 0), 2),
          round(e.get("reference_gpu_kernel", {}).get("value", 0), 1)))
print("ref kernels", {k: (round(v["value"], 1), v.get("tuned")) for k, v in d["reference_gpu_kernels"].items()})
PY
tail -3 $O/bench.err
