#!/bin/bash
# A/B of Jacobian builds: bench with the Jacobian leg only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/abj.log
for rep in 1 2; do
for v in "$@"; do
  if [ $v = new ]; then lib=structured_b200/libstructured_gpu.so; else lib=scratch/variants/$v.so; fi
  SGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-linsolve 2>>gpurun_out/abj.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['roofline']['kernel_ms'], d['jacobian'])
" >> gpurun_out/abj.log
done
done
cat gpurun_out/abj.log
