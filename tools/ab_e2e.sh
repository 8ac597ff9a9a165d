#!/bin/bash
# A/B of the host-buffer pipeline (sgpu_residual_host): e2e Mcell/s for chunk counts / ramped chunk sizes
# usage: tools/ab_e2e.sh "RAMP CHUNKS" ...   e.g. tools/ab_e2e.sh "0 18" "1 18" "1 23"
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; : > gpurun_out/ab_e2e.log
for rep in 1 2; do
for cfg in "$@"; do
  r=${cfg% *}; n=${cfg#* }
  SGPU_PIPE_RAMP=$r SGPU_PIPE_CHUNKS=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-jacobian --no-linsolve 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('ramp/chunks $cfg', e['value'], e.get('copy_ceiling',{}).get('value'), e.get('rel_diff_vs_device_path'))
" >> gpurun_out/ab_e2e.log
done
done
cat gpurun_out/ab_e2e.log
