#!/bin/bash
# Round 2: A/B numbers of the data-reuse mode (--fuse reuse) against the default sweep kernels, at the c1 / c4 / c2 sizes.
O=gpurun_out/r02_call13
mkdir -p $O
export PROBE_SWEEPS=40
{
python tools/probe_shape.py 2d5pt_star 4096,4096 '{}' '{"fuse":"reuse"}' '{"fuse":"reuse","bx":256,"sn":64}' '{"fuse":"reuse","bx":64,"sn":16}'
python tools/probe_shape.py 3d7pt_star 768,768,768 '{}' '{"fuse":"reuse"}' '{"fuse":"reuse","bx":64,"by":8,"sn":64}' '{"fuse":"reuse","bx":32,"by":16,"sn":16}'
python tools/probe_shape.py 2d9pt_box 16384,16384 '{"step":4}' '{"step":4,"fuse":"algebraic"}' '{"step":4,"fuse":"reuse"}' '{"step":4,"fuse":"reuse","bx":256,"sn":64}'
} 2>&1 | tee $O/reuse_ab.txt | cut -c1-190
