#!/bin/bash
# GPU run 1 of round 2: full GPU test suite, the bench the way the driver runs it, the default bench, the reference arm
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q --durations=15) > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests1.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_20.json 2> gpurun_out/r2_bench_20.err
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2>&1
tail -5 gpurun_out/r2_tests1.log; nproc; free -g | head -2
