#!/bin/bash
# One gpurun call that produces the round's single-GPU evidence under gpurun_out/ (then: python tools/make_profiles.py <tag>).
#   GPU tests, smoke, the contract bench (both arms), ncu --set full of the fused 3D temporal kernel,
#   and the ncu launch list (+ DRAM bytes) of the bench command.  Every ncu pass runs only after the same
#   command has exited 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
t0=$(date +%s)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
t1=$(date +%s)
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? seconds=$(( $(date +%s) - t1 ))" >> gpurun_out/bench.err
t2=$(date +%s)
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm rc=$? seconds=$(( $(date +%s) - t2 ))" >> gpurun_out/bench_reference.err
for p in c4t2 c1; do
  python -m drstencil_b200.tuner.run_one --preset $p > gpurun_out/plain_$p.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:dr_ -s 3 -c 2 -f -o gpurun_out/prof_$p \
      python -m drstencil_b200.tuner.run_one --preset $p > gpurun_out/ncu_$p.log 2>&1
done
python bench.py --no-extras --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
cp gpurun_out/launches.csv gpurun_out/dram_c5.csv
echo "total seconds $(( $(date +%s) - t0 ))"
tail -2 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -1 gpurun_out/bench.err; tail -1 gpurun_out/bench_reference.err
