#!/bin/bash
# Round 2, final single-GPU sanity: the whole GPU suite and smoke() on the final commit.
O=gpurun_out/r02_call11
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; cat $O/smoke.log | tail -4
