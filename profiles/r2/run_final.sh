#!/bin/bash
# Round-2 records: full GPU suite, driver-style short bench, default bench (with secondary configs), reference arm,
# ncu launch list + one --set full capture of the step kernel (mid-day and night)
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_20.json 2> gpurun_out/r2_bench_20.err
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 40 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
PROF_STEPS=152 ncu --set full --clock-control none --import-source on -k regex:ev_step_main -s 150 -c 1 -f -o gpurun_out/r2_prof_ev_midday \
    python profiles/prof_ev.py > gpurun_out/r2_prof_ev_midday.log 2>&1
PROF_STEPS=14 ncu --set full --clock-control none --import-source on -k regex:ev_step_main -s 12 -c 1 -f -o gpurun_out/r2_prof_ev_night \
    python profiles/prof_ev.py > gpurun_out/r2_prof_ev_night.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
ncu --set full --clock-control none --import-source on -k regex:building_step1 -s 30 -c 1 -f -o gpurun_out/r2_prof_building_step1 \
    python profiles/r2/building_step_probe.py > gpurun_out/r2_prof_building_step1.log 2>&1
tail -4 gpurun_out/r2_tests.log; tail -2 gpurun_out/r2_smoke.log
