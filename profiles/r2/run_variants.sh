#!/bin/bash
# times the EV step (bench.py headline only) with each library variant given as argument ("default" = the in-tree .so)
mkdir -p gpurun_out
for L in "$@"; do
  if [ "$L" = "default" ]; then unset SG_LIB_PATH; else export SG_LIB_PATH=$PWD/$L; fi
  python bench.py --steps ${VSTEPS:-4000} --warmup 400 --no-cpu-baseline --no-secondary > gpurun_out/var.json 2> gpurun_out/var.err || tail -5 gpurun_out/var.err
  python - "$L" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open("gpurun_out/var.json") if l.startswith("{")][-1])
    print(f"{sys.argv[1]:40s} us/step {d['ms_per_step']*1e3:7.2f} frac {d['roofline']['frac']:.3f} pairs {d['roofline']['kernel_ms_event_pairs_whole_episode']*1e3:7.2f} e2e {d['e2e']['value']:.3e} nproj/step {d['config']['n_projected_per_step']:.1f} mean_r {d['reward_stats']['mean']:.6f}", flush=True)
except Exception as e:
    print(sys.argv[1], "FAILED", e, flush=True)
PY
done 2>&1 | tee -a gpurun_out/variants.log
