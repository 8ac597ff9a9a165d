#!/bin/bash
# A/B of library builds on one explicit rk4 step (bench.py field explicit_step)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/ab_explicit.log
for rep in 1 2; do
for v in "$@"; do
  if [ $v = new ]; then lib=structured_b200/libstructured_gpu.so; else lib=scratch/variants/$v.so; fi
  SGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-linsolve 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['explicit_step']['rk4_step_ms'], d['explicit_step']['rk4_step_two_kernel_ms'], d['roofline']['kernel_ms'])
" >> gpurun_out/ab_explicit.log
done
done
cat gpurun_out/ab_explicit.log
