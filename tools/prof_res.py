#!/usr/bin/env python
"""Five residual evaluations of the SA bump-channel workload -- the command `ncu -k regex:residual_kernel` wraps
(tools/prof_res.py [nic njc]; SGPU_LIB selects the library build)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from structured_b200.api import GpuEulerEquation
from structured_b200.cases import turbulent_channel_case

nic, njc = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
case = turbulent_channel_case(nic, njc, ntrans=1)
eq = GpuEulerEquation(case)
eq.set_state(case.perturbed_q())
for _ in range(5):
    eq.residual_device(0)
eq.synchronize()
eq.close()
print("done")
