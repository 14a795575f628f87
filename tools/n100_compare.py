import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from tinympc_b200 import abi, workloads as wl
from tinympc_b200.solver import BatchedTinySolver, setup_problem
for (nx, nu, N) in ((12, 4, 100), (4, 2, 100), (16, 8, 100), (16, 8, 50), (12, 4, 50)):
    spec = wl.random_lti(nx, nu, N, seed=1); spec.settings.abs_pri_tol = 0.0; spec.settings.abs_dua_tol = 0.0; spec.settings.max_iter = 50
    inst = wl.random_instances(131072, nx, N, seed=2)
    prob = setup_problem(spec, np.float32)
    for kname, k in (("gpi", abi.KERNEL_GPI), ("tpi", abi.KERNEL_TPI), ("auto", abi.KERNEL_AUTO)):
        s = BatchedTinySolver(prob, spec.settings, kernel=k)
        batch, out = s.make_device_batch(inst["x0"], inst["Xref"], None, cold_start=True)
        ms = []
        for _ in range(3):
            s.solve_device(batch); torch.cuda.synchronize(); ms.append(s.stats()["kernel_ms"])
        st = s.stats()
        print(f"({nx},{nu},{N}) {kname}->{st['kernel_family']} L={st['lanes_per_instance']} thr={st['threads_per_cta']}: {min(ms[1:]):.2f} ms", flush=True)
        s.close()
