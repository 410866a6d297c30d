#!/bin/bash
# EV step time for several tilings of the 65 536-env batch (SG_EV_TILING="t_full,tail_envs"; unset = the launcher's choice)
for t in ${TILINGS:-"" "4096,16" "3520,6" "3520,8" "0,12" "0,8"}; do
  if [ -z "$t" ]; then unset SG_EV_TILING; else export SG_EV_TILING=$t; fi
  python bench.py --steps 2000 --warmup 300 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print('tiling', os.environ.get('SG_EV_TILING','(policy)'), 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], 'mean_r %.6f'%d['reward_stats']['mean'])"
done
