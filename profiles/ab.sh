#!/bin/bash
# A/B helper: GPU tests, then bench with each library variant given as arguments.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E|passed|failed|Error" | tail -12 | tee gpurun_out/ab_tests.log
for L in "$@"; do
  echo "LIB=$L"
  if [ "$L" = "default" ]; then unset SG_LIB_PATH; else export SG_LIB_PATH=$L; fi
  python bench.py --steps 8000 --warmup 1200 --no-cpu-baseline > gpurun_out/ab_bench.log 2>&1
  tail -3 gpurun_out/ab_bench.log | cut -c1-900
done
