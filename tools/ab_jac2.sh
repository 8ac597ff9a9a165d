#!/bin/bash
# Jacobian build time of the in-tree library: single-pass marching kernel vs the round-1 two-stage build (SGPU_JAC=two_stage)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/abj2.log
for rep in 1 2; do
for mode in march two_stage; do
  SGPU_JAC=$mode timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-linsolve 2>>gpurun_out/abj2.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$mode', d['jacobian'])
" >> gpurun_out/abj2.log
done
done
cat gpurun_out/abj2.log
