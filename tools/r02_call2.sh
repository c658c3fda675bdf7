#!/bin/bash
# Round 2, GPU call 2 (2 GPUs): the fused slab protocol -- parity (Python and plain-C drivers), N=2 bench lines.
O=gpurun_out/r02_call2
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/slab_check.py > $O/slab_check.txt 2>&1; echo "slab_check rc=$?" >> $O/slab_check.txt
grep -E "slab_check|SLAB_CHECK|rc=|Error|error" $O/slab_check.txt | tail -60
timeout 600 python -m pytest tests/test_c_consumer.py -m gpu -x -q > $O/pytest_c_consumer.txt 2>&1; echo "rc=$?" >> $O/pytest_c_consumer.txt
tail -5 $O/pytest_c_consumer.txt
timeout 900 $TR --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
timeout 600 $TR --master-port 29622 bench.py --gpus 2 --steps 5 --warmup 3 --halo p2p-flags --no-extras --no-parity > $O/bench_n2_flags.json 2> $O/bench_n2_flags.err; echo "bench n2 flags rc=$?"
timeout 600 $TR --master-port 29623 bench.py --gpus 2 --steps 5 --warmup 3 --halo nccl --no-extras --no-parity > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
python - <<'PY'
import json
for f in ("bench_n2", "bench_n2_flags", "bench_n2_nccl"):
    try:
        d = json.load(open("gpurun_out/r02_call2/%s.json" % f))
        print(f, "value %.1f ms/step %.1f launches %d frac %.3f e2e %.1f (%s)" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["api"][:40]))
        print("   parity", json.dumps(d.get("parity"))[:600])
        print("   bound", d["e2e"].get("bound"), "cpu", (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 $O/bench_n2.err
