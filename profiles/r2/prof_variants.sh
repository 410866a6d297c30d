#!/bin/bash
# ncu --set full of one mid-day launch (t = 150) of the EV step kernel for each library variant given
mkdir -p gpurun_out
for L in "$@"; do
  tag=$(basename "$L" .so)
  if [ "$L" = "default" ]; then unset SG_LIB_PATH; tag=default; else export SG_LIB_PATH=$PWD/$L; fi
  PROF_STEPS=${PROF_STEPS:-152} ncu --set full --clock-control none --import-source on -k regex:ev_step_main -s ${PROF_SKIP:-150} -c 1 \
      -f -o gpurun_out/prof_ev_$tag python profiles/prof_ev.py > gpurun_out/prof_ev_$tag.log 2>&1
  tail -2 gpurun_out/prof_ev_$tag.log
done
