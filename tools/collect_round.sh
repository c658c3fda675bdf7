#!/bin/bash
# One gpurun call that produces the round's single-GPU evidence under gpurun_out/ (then: python tools/make_profiles.py <tag>).
#   GPU tests, smoke, the contract bench (both arms), ncu --set full of the fused 3D temporal kernel,
#   and the ncu launch list (+ DRAM bytes) of the bench command.  Every ncu pass runs only after the same
#   command has exited 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
t0=$(date +%s)
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
t1=$(date +%s)
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? seconds=$(( $(date +%s) - t1 ))" >> gpurun_out/bench.err
t2=$(date +%s)
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm rc=$? seconds=$(( $(date +%s) - t2 ))" >> gpurun_out/bench_reference.err
for p in c1 c2 c3 c4 c4t2; do
  python -m drstencil_b200.tuner.run_one --preset $p > gpurun_out/plain_$p.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dr_ -s 3 -c 2 -f -o gpurun_out/prof_$p \
      python -m drstencil_b200.tuner.run_one --preset $p > gpurun_out/ncu_$p.log 2>&1
done
# the headline kernel (c5 preset) on a 256-plane slab of the c5 grid: 54 GiB of state is too much for --set full replays
C5="--sn 64 --rows-3d 8 --share-x 2 --share-y 2"
python -m drstencil_b200.tuner.run_one stc/3d7pt_star.stc --3d --size 256 1536 1536 -- $C5 > gpurun_out/plain_c5slab.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dr_ -s 3 -c 2 -f -o gpurun_out/prof_c5slab \
    python -m drstencil_b200.tuner.run_one stc/3d7pt_star.stc --3d --size 256 1536 1536 -- $C5 > gpurun_out/ncu_c5slab.log 2>&1
python bench.py --no-extras --no-parity --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --no-extras --no-parity --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
cp gpurun_out/launches.csv gpurun_out/dram_c5.csv
echo "total seconds $(( $(date +%s) - t0 ))"
tail -2 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -1 gpurun_out/bench.err; tail -1 gpurun_out/bench_reference.err
