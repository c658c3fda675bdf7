"""Auto-tuner of the B200 stencil engine -- the rebuild of the reference's tuning flow
(/root/reference/benchmarks/*/{starter.sh,tuning.py,compile_run.sh,getGpuMetrics.{sh,py}}).

  space.py    the search space (the reference's axes plus the engine's own), its validity filter
              and the reference's configuration-name grammar (tuning.py:72-86)
  tune.py     model-pruned search: candidates are timed in-process with CUDA events through the
              C ABI (no per-candidate nvcc + process launch), the winners are then justified with
              Nsight Compute metrics collected BY NAME
  metrics.py  `ncu --csv` parser keyed on metric names (replaces the order-dependent scrape of
              getGpuMetrics.py:9, which only understands Nsight Compute 2020.3 output)
  run_one.py  launches one configuration a few times -- the process ncu wraps
"""
