#!/usr/bin/env python3
"""Developer micro-benchmark (not the contract bench): device-resident timing of one batched solve.
usage: python tools/quick_bench.py [--kernel tpi|gpi|auto] [--mode strict|fast] [--config c2|c3|c4] [--B n] [--reps n]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from tinympc_b200 import abi, workloads as wl  # noqa: E402
from tinympc_b200.solver import BatchedTinySolver, setup_problem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default="auto")
ap.add_argument("--mode", default="strict")
ap.add_argument("--config", default="c2")
ap.add_argument("--B", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--max_iter", type=int, default=0)
a = ap.parse_args()
K = dict(auto=abi.KERNEL_AUTO, tpi=abi.KERNEL_TPI, gpi=abi.KERNEL_GPI, gps=abi.KERNEL_GPS)[a.kernel]
M = dict(strict=abi.MODE_STRICT, fast=abi.MODE_FAST)[a.mode]
if a.config == "c2":
    spec, dt, B = wl.quadrotor(N=50), np.float32, a.B or 65536
    inst = wl.hovering_instances(B, N=50, dtype=dt)
elif a.config == "c3":
    spec, dt, B = wl.quadrotor(N=50), np.float32, a.B or 65536
    inst = wl.tracking_instances(B, N=50, seed=0, dtype=dt)
elif a.config in ("c4", "c4p"):  # c4p: one reference trajectory per instance (what bench.py's C4 entry runs)
    spec, dt, B = wl.rocket(N=100), np.float64, a.B or 16384
    inst = wl.rocket_instances(B, N=100, seed=0, dtype=dt, per_instance_refs=a.config == "c4p")
else:
    raise SystemExit("config")
if a.max_iter:
    spec.settings.max_iter = a.max_iter
prob = setup_problem(spec, dt)
s = BatchedTinySolver(prob, spec.settings, device=0, mode=M, kernel=K)
batch, out = s.make_device_batch(inst["x0"], inst["Xref"], inst.get("Uref"), cold_start=True)
ms = []
for r in range(a.reps + 2):
    s.solve_device(batch)
    torch.cuda.synchronize()
    ms.append(s.stats()["kernel_ms"])
st = s.stats()
iters = int(out["iter"].sum().item())
best = min(ms[2:] or ms)
print(f"{a.config} kernel={a.kernel}->{st['kernel_family']} mode={a.mode} B={B} iters={iters} solved={int(out['solved'].sum().item())} "
      f"ms(all)={[round(m, 3) for m in ms]} best={best:.3f} ms  -> {B / best * 1e3:.3e} inst/s  {iters / best * 1e3:.3e} ADMM it/s "
      f"ws={st["workspace_bytes"] >> 20}MiB gpiB={st['gpi_instances']} ctas={st['ctas']} thr={st['threads_per_cta']} smem={st['smem_bytes_per_cta']} L={st['lanes_per_instance']}")
