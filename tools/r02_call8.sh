#!/bin/bash
# Round 2, 8-GPU re-measurement after the protocol change: the contract bench (parity included) and the separate-flag-kernel A/B.
N=8
O=gpurun_out/r02_p2_n8
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 400 $TR --master-port 29622 bench.py --gpus $N --steps 10 --warmup 3 --halo p2p-flags --no-extras --no-parity --e2e-plain > $O/bench_flags.json 2> $O/bench_flags.err; echo "bench flags rc=$?"
python - <<PY
import json
for f in ("bench", "bench_flags"):
    try:
        d = json.load(open("$O/%s.json" % f))
    except Exception as e:
        print(f, "unreadable:", str(e)[:80]); continue
    print(f, "value %.1f ms/step %.2f launches %d frac %.3f clocks %s | e2e %.1f ms/step %.1f (%s)" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["api"][:30]))
    if d.get("parity"):
        print("   parity ok=%s host_ok=%s" % (d["parity"].get("ok"), d["parity"].get("host_path_ok")))
    if d.get("temporal_fused"):
        print("   depth 2:", d["temporal_fused"].get("value"))
PY
