#!/bin/bash
# A/B of two builds of libmarl_b200.so on the headline bench:  tools/bench_ab.sh gpurun_ab/lib_base.so gpurun_ab/lib_new.so
for lib in "$@"; do
  MARL_AB_LIB=$lib timeout -s KILL 400 python -c "
import sys, runpy
sys.path.insert(0, 'tools'); import ab
sys.argv = ['bench.py', '--steps', '5', '--warmup', '3']
runpy.run_path('bench.py', run_name='__main__')" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', 'value %.2f M  ms %.2f  e2e %.2f M  env_only %.2f ms  train %.1f ms' % (d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6, d['env_only']['ms_per_episode'], d['train']['ms_per_epoch']))"
done
