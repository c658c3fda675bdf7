#!/bin/bash
# Round 2, 2-GPU check after the index-arithmetic change in the fused 3D temporal kernel: slab parity + depth-2 bench.
O=gpurun_out/r02_call12
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
DRS_SLAB_CHECK_HOST=0 timeout 600 $TR --master-port 29611 tools/slab_check.py > $O/slab_check.txt 2>&1; echo "slab_check rc=$?"; grep -cE "bit-exact" $O/slab_check.txt; grep -E "MISMATCH|SLAB_CHECK" $O/slab_check.txt | head -5
timeout 600 $TR --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 --depth 2 --no-extras --e2e-plain > $O/bench_depth2.json 2> $O/bench_depth2.err; echo "bench depth2 rc=$?"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "3d and (temporal or whole)" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_call12/bench_depth2.json"))
print("depth 2, N=2: %.1f GStencil/s, %.1f ms/step, launches %d, parity %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["parity"]["ok"] if d.get("parity") else None))
PY
