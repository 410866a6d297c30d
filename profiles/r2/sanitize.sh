#!/bin/bash
# compute-sanitizer racecheck + memcheck on ev_step_main_kernel, building_mma_kernel, cogen_mlp_tc_kernel
mkdir -p gpurun_out
for target in ${SAN_TARGETS:-ev building building_step market cogen}; do
  for tool in memcheck racecheck; do
    echo "==== $target / $tool"
    timeout 420 compute-sanitizer --tool $tool --print-limit 8 python profiles/r2/sanitize_targets.py $target 2>&1 | grep -vE "^\s*$" | tail -14
  done
done 2>&1 | tee gpurun_out/r2_sanitizer.log
