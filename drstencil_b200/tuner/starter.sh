#!/bin/bash
# Tuning entry point, same role as the reference's benchmarks/<stencil>/starter.sh:
#   ./starter.sh <stc file> [tune.py options]        e.g.  ./starter.sh ../../stc/2d9pt_box.stc --step 4 --ncu
# Writes tuning_result.json, duration.log and tuning-time.log in the current directory.
set -e
start=$(date +%s)
python -m drstencil_b200.tuner.tune "$@"
end=$(date +%s)
echo "tuning time: $((end - start)) s" >> tuning-time.log
