#!/bin/bash
# usage: tools/sass_loop_stats.sh [extra nvcc flags]   (no GPU needed)
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
W=${TMPDIR:-/tmp}/sg_sass; mkdir -p $W
cat > $W/rk.cu <<EOC
#include "$HERE/include/structured_gpu.h"
#include "common.cuh"
#include "residual_kernel.cuh"
template __global__ void sg::residual_kernel<5, 2, SGPU_FLUX_ROE, true, SG_SASS_UPD>(const sg::ResParams);
EOC
nvcc -DSG_SASS_UPD=${SG_SASS_UPD:-false} -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I$HERE/structured_b200/csrc "$@" -Xptxas -v -cubin -o $W/rk.cubin $W/rk.cu 2>&1 | grep -E "registers|spill|error"
cuobjdump -sass $W/rk.cubin > $W/rk.sass
python3 - $W/rk.sass <<'PY'
import re, collections, sys
L = open(sys.argv[1]).read().split('\n')
ins = []
i = 0
while i < len(L):
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+).*?/\* (0x[0-9a-f]{16}) \*/', L[i])
    if m and i + 1 < len(L):
        m2 = re.search(r'/\* (0x[0-9a-f]{16}) \*/', L[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), m.group(3), (hi >> 41) & 0xf, L[i]))   # stall count = bits 41..44 of the control word
            i += 2
            continue
    i += 1
loops = []                                      # backward branches = loops
for a, op, st, l in ins:
    if op.startswith('BRA'):
        t = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', l)
        if t:
            tgt = int(t.group(1), 16)
            if tgt < a: loops.append((tgt, a))
loops.sort(key=lambda x: x[0] - x[1])           # largest first
best = loops[0]                                 # the row loop = the largest loop that is not a wrapper around another big loop
for cand in loops:                              # (the segment loop of the balanced partition contains the row loop)
    inner = [x for x in loops if x != cand and cand[0] <= x[0] and x[1] <= cand[1] and (x[1] - x[0]) > 0.4*(cand[1] - cand[0])]
    if not inner:
        best = cand
        break
loop = [x for x in ins if best[0] <= x[0] <= best[1]]
c = collections.Counter(x[1].split('.')[0] for x in loop)
fp64 = sum(v for k, v in c.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
print('row loop: %d instructions, fp64 %d, sum of stall counts %d' % (len(loop), fp64, sum(x[2] for x in loop)))
print({k: c[k] for k in ('DFMA', 'DMUL', 'DADD', 'MUFU', 'LDS', 'STS', 'LDGSTS', 'IMAD', 'LDL', 'STL', 'BRA', 'BSSY')})
PY
