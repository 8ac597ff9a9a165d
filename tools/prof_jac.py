#!/usr/bin/env python
"""One Jacobian build of the SA channel workload at a given size (for `ncu -k regex:jac_march`): tools/prof_jac.py [nic njc]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from structured_b200.api import GpuEulerEquation
from structured_b200.cases import turbulent_channel_case

nic, njc = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
case = turbulent_channel_case(nic, njc, ntrans=1)
eq = GpuEulerEquation(case)
eq.set_state(case.perturbed_q())
for _ in range(3):
    slots, ms = eq.jacobian_device()
    print("jacobian build %.3f ms (%d slots)" % (ms, slots))
eq.close()
