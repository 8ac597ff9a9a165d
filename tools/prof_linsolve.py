#!/usr/bin/env python
"""One short GMRES solve with the j-line preconditioner on the SA channel workload (for `ncu -k regex:line_factor` etc.):
tools/prof_linsolve.py [nic njc]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from structured_b200.api import GpuEulerEquation
from structured_b200.cases import turbulent_channel_case

nic, njc = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 1024)
case = turbulent_channel_case(nic, njc, ntrans=1)
eq = GpuEulerEquation(case)
eq.set_state(case.perturbed_q())
eq.calc_dt(20.0)
eq.residual_device(0)
eq.jacobian_device()
x, info = eq.linear_solve("lhs", precond="line_j", restart=20, max_iter=20, rtol=1e-12)
print(info)
eq.close()
