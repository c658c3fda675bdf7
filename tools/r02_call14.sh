#!/bin/bash
# Round 2: tile shapes of the fused 3D temporal kernel (depth 2) on c5 after the deferred-scale change (registers 128 -> 122).
O=gpurun_out/r02_call14
mkdir -p $O
export PROBE_SWEEPS=40
python tools/probe_shape.py 3d7pt_star 1536,1536,1536 '{"step":2}' '{"step":2,"warps":16,"rows_3d":4,"min_blocks":1}' '{"step":2,"stages":4}' \
   '{"step":2,"warps":16,"rows_3d":2}' '{"step":2,"warps":12,"rows_3d":4,"min_blocks":1}' '{"step":2,"sn":256}' '{"step":2,"sn":64}' \
   '{"step":2,"warps":16,"rows_3d":4,"min_blocks":1,"sn":256}' 2>&1 | tee $O/t3_shapes.txt | cut -c1-200
