#!/bin/bash
# A/B of residual-kernel builds on one box: same bench command, SGPU_LIB selects the library
# usage: ab.sh variant...   ("new" = the in-tree library)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do
for v in "$@"; do
  if [ $v = new ]; then lib=structured_b200/libstructured_gpu.so; else lib=scratch/variants/$v.so; fi
  SGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-jacobian --no-linsolve 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
" >> gpurun_out/ab.log
done
done
cat gpurun_out/ab.log
