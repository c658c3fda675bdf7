#!/bin/bash
# Round 2, GPU call 1 (one GPU): time the CTA-shared input ring of the single-step 3D sweep on c5
# (sustained: 60 sweeps per sample), then ncu --set full of the default and of the shared-ring
# kernel on a 256-plane c5 slab.  Output under gpurun_out/r02_call1/.
O=gpurun_out/r02_call1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
lscpu | head -25 > $O/lscpu.txt 2>&1; nproc >> $O/lscpu.txt
numactl -H >> $O/lscpu.txt 2>&1
DRS_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_shared_ring.py -m gpu -x -q > $O/pytest_shared_ring.log 2>&1
echo "pytest rc=$?" >> $O/pytest_shared_ring.log
export PROBE_SWEEPS=60
timeout 900 python tools/probe_shape.py 3d7pt_star 1536,1536,1536 \
  '{"sn":64,"rows_3d":6,"warps":4}' \
  '{"sn":64,"rows_3d":6,"share_x":2,"share_y":2}' \
  '{"sn":64,"rows_3d":4,"share_x":2,"share_y":2}' \
  '{"sn":64,"rows_3d":8,"share_x":2,"share_y":2}' \
  '{"sn":64,"rows_3d":6,"share_x":3,"share_y":1}' \
  '{"sn":64,"rows_3d":6,"share_x":3,"share_y":2}' \
  '{"sn":64,"rows_3d":4,"share_x":3,"share_y":2}' \
  '{"sn":64,"rows_3d":4,"share_x":3,"share_y":3}' \
  '{"sn":64,"rows_3d":6,"share_x":2,"share_y":3}' \
  '{"sn":64,"rows_3d":4,"share_x":2,"share_y":4}' \
  '{"sn":64,"rows_3d":6,"share_x":1,"share_y":4}' \
  '{"sn":64,"rows_3d":6,"share_x":2,"share_y":1}' \
  '{"sn":128,"rows_3d":6,"share_x":2,"share_y":2}' \
  '{"sn":64,"rows_3d":6,"share_x":2,"share_y":2,"stages":8}' \
  '{"sn":64,"rows_3d":6,"share_x":3,"share_y":2,"stages":8}' \
  > $O/probe_c5_share.txt 2>&1
echo "probe rc=$?" >> $O/probe_c5_share.txt
# c4 too (768^3): does the shared ring cost anything where planes fit the L2?
timeout 300 python tools/probe_shape.py 3d7pt_star 768,768,768 \
  '{"sn":16,"rows_3d":4}' '{"sn":16,"rows_3d":4,"share_x":2,"share_y":2}' '{"sn":64,"rows_3d":6,"share_x":2,"share_y":2}' \
  '{"sn":16,"rows_3d":6,"share_x":3,"share_y":2}' > $O/probe_c4_share.txt 2>&1
# ncu --set full on a 256-plane c5 slab: default kernel and shared ring (each only after a plain run exited 0)
for v in base share22 share32; do
  case $v in
    base) OPTS="--sn 64 --rows-3d 6 --warps 4";;
    share22) OPTS="--sn 64 --rows-3d 6 --share-x 2 --share-y 2";;
    share32) OPTS="--sn 64 --rows-3d 6 --share-x 3 --share-y 2";;
  esac
  python -m drstencil_b200.tuner.run_one stc/3d7pt_star.stc --3d --size 256 1536 1536 -- $OPTS > $O/plain_$v.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dr_ -s 3 -c 2 -f -o $O/prof_c5slab_$v \
    python -m drstencil_b200.tuner.run_one stc/3d7pt_star.stc --3d --size 256 1536 1536 -- $OPTS > $O/ncu_$v.log 2>&1
done
cat $O/probe_c5_share.txt | cut -c1-200
tail -3 $O/pytest_shared_ring.log
