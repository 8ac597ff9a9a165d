#!/bin/bash
# A/B of library builds on the device GMRES sample of bench.py (matvec / line sweep / setup / per-iteration times)
# usage: tools/ab_linsolve.sh variant...   ("new" = the in-tree library, others = scratch/variants/<name>.so)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/ab_linsolve.log
for rep in 1 2; do
for v in "$@"; do
  if [ $v = new ]; then lib=structured_b200/libstructured_gpu.so; else lib=scratch/variants/$v.so; fi
  SGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)['linear_solve']; print('$v', 'iter', d['ms_per_iteration'], 'matvec', d['matvec_ms'], 'precond', d['precond_ms'], 'setup', d['setup_ms'], 'rel', d['rel_residual'])
" >> gpurun_out/ab_linsolve.log
done
done
cat gpurun_out/ab_linsolve.log
