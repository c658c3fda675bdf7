#!/bin/bash
# Round 2, multi-GPU call: bash tools/r02_call4.sh <N>  -- the contract bench at N GPUs (parity object included), the
# A/B of the exchange protocols, the host-link probe, and (N <= 4) the slab tests.
N=$1
O=gpurun_out/r02_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
t0=$(date +%s)
timeout 900 $TR --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench n$N rc=$? seconds=$(( $(date +%s) - t0 ))"
if [ "$N" = "8" ] || [ -n "$AB" ]; then
  timeout 600 $TR --master-port 29622 bench.py --gpus $N --steps 10 --warmup 3 --halo p2p-flags --no-extras --no-parity --e2e-plain > $O/bench_flags.json 2> $O/bench_flags.err; echo "bench flags rc=$?"
  timeout 600 $TR --master-port 29623 bench.py --gpus $N --steps 10 --warmup 3 --halo nccl --no-extras --no-parity --e2e-plain > $O/bench_nccl.json 2> $O/bench_nccl.err; echo "bench nccl rc=$?"
fi
timeout 300 $TR --master-port 29624 tools/hostlink_probe.py 3 > $O/hostlink.txt 2>&1; echo "hostlink rc=$?"
cp gpurun_out/hostlink_n$N.json $O/ 2>/dev/null
grep -E "H2D|D2H" $O/hostlink.txt
if [ "$N" -le 4 ]; then
  timeout 900 python -m pytest tests/test_slab_gpu.py tests/test_c_consumer.py -m gpu -q > $O/pytest_slab.txt 2>&1; echo "pytest slab rc=$?"; tail -3 $O/pytest_slab.txt
fi
nvidia-smi topo -m > $O/topo.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name" > $O/lscpu.txt
python - <<PY
import json
for f in ("bench", "bench_flags", "bench_nccl"):
    try:
        d = json.load(open("$O/%s.json" % f))
    except Exception as e:
        print(f, "unreadable:", str(e)[:80]); continue
    print(f, "value %.1f ms/step %.1f launches %d frac %.3f | e2e %.1f ms/step %.1f (%s)" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["api"][:36]))
    print("   bound", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].get("bound", {}).items() if k != "note"})
    if d.get("parity"):
        print("   parity ok=%s host_ok=%s" % (d["parity"].get("ok"), d["parity"].get("host_path_ok")), [(c["case"][:12], c["slab_vs_single"], c.get("host_path_vs_single")) for c in d["parity"]["cases"]])
    if d.get("cpu_baseline"):
        print("   cpu", d["cpu_baseline"]["cores"], "cores", round(d["cpu_baseline"]["value"], 2), "GStencil/s")
    if d.get("temporal_fused"):
        print("   depth 2:", d["temporal_fused"])
PY
tail -3 $O/bench.err
echo "total seconds $(( $(date +%s) - t0 ))"
