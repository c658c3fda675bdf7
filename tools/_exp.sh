for d in "DRS_T3_OWNREG=3" "DRS_T3_OWNREG=4"; do
  echo "== defines: [$d]"
  DRS_EXTRA_DEFINES="$d" python -m pytest tests/test_parity_gpu.py -q -x -k "3d_temporal_depth" 2>&1 | tail -1
  DRS_EXTRA_DEFINES="$d" python tools/probe_shape.py 3d7pt_star 768,768,768 '{"step":2}' '{"step":2,"stages":4}' '{"step":3}' 2>&1 | tail -3
  DRS_EXTRA_DEFINES="$d" python tools/probe_shape.py 3d7pt_star 1536,1536,1536 '{"step":2}' '{"step":2,"stages":4}' 2>&1 | tail -2
done
