#!/usr/bin/env python3
"""Developer benchmark: device-resident closed-loop MPC (DeviceMPCLoop), B plants tracking the y-axis line, N=50, fp32."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from tinympc_b200 import abi, workloads as wl
from tinympc_b200.closed_loop import DeviceMPCLoop
from tinympc_b200.solver import BatchedTinySolver, setup_problem

B, N, STEPS = 65536, 50, 12
spec = wl.quadrotor(N=N)
prob = setup_problem(spec, np.float32)
inst = wl.tracking_instances(B, N=N, seed=0, dtype=np.float32)
Xref = torch.as_tensor(inst["Xref"], device="cuda")
for kname, k in (("gpi", abi.KERNEL_GPI), ("tpi", abi.KERNEL_TPI)):
    for exact, sol in ((True, True), (False, True), (False, False)):
        s = BatchedTinySolver(prob, spec.settings, kernel=k)
        loop = DeviceMPCLoop(s, inst["x0"], reset_duals=True, exact_first_residual=exact)
        loop.want_solution = sol
        fields = loop.fields
        ts, its = [], []
        for step in range(STEPS):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = loop.step(Xref)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
            its.append(float(out["iter"].float().mean().item()))
        print(f"{kname} persist={'v,z' if 'v' in fields else 'no v,z'} sol={sol}: ms/step {[round(t,2) for t in ts]}  mean iters {[round(i,2) for i in its]}  solved {int(out['solved'].sum().item())}/{B}", flush=True)
        s.close()
