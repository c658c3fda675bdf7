#!/bin/bash
# Round 2: input-ring depth of the fused 3D temporal kernel (depth 2) on c5 -- 2 (default), 3, 4, 5 stages, resident CTAs per SM
# as the driver reports them, and the same with an explicit shared-memory carve-out preference.
O=gpurun_out/r02_call17
mkdir -p $O
export PROBE_SWEEPS=40 DRS_DEBUG_OCC=1
{
echo "== carve-out: driver default"
timeout 200 python tools/probe_shape.py 3d7pt_star 1536,1536,1536 '{"step":2}' '{"step":2,"stages":3}' '{"step":2,"stages":4}' '{"step":2,"stages":5}' '{"step":2}' '{"step":2,"stages":3}'
echo "== carve-out: DRS_CARVEOUT=100"
DRS_CARVEOUT=100 timeout 200 python tools/probe_shape.py 3d7pt_star 1536,1536,1536 '{"step":2}' '{"step":2,"stages":3}' '{"step":2,"stages":4}'
} 2>&1 | tee $O/t3_ring_depth.txt | cut -c1-220
