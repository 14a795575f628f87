#!/usr/bin/env python3
"""Generate the golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

For each parity case (tests/helpers.py:make_cases) this runs the reference — compiled with the pinned flags
into oracle/_ref/libtinympc_ref_{f64,f32}.so by oracle/Makefile — through a warm-started closed loop and stores,
per case, one compressed .npz with
  * the complete problem (model, the cache the reference's tiny_setup computed, bounds, cones, hyperplanes),
  * the settings, the inputs of every step (x0 sequence, Xref, Uref),
  * the reference's outputs of every step (solution, iter, solved, residuals, and every state array).
The fixtures pin the oracle restatement (tests/test_oracle_golden.py, runs anywhere) and the CUDA path
(tests/test_gpu_parity.py) to the reference bit-for-bit without /root/reference being present.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import helpers as H  # noqa: E402
from oracle import oracle  # noqa: E402
from tinympc_b200 import abi  # noqa: E402

PROBLEM_FIELDS = ["A", "B", "f", "Q", "R", "Kinf", "Pinf", "Quu_inv", "AmBKt", "APf", "BPf", "x_min", "x_max", "u_min",
                  "u_max", "Acx", "qcx", "cx", "Acu", "qcu", "cu", "Alin_x", "blin_x", "Alin_u", "blin_u", "tv_Alin_x",
                  "tv_blin_x", "tv_Alin_u", "tv_blin_u"]


def main():
    cases = H.make_cases()
    for name in sorted(cases):
        c = cases[name]
        prob = H.problem_from_spec(c["spec"], c["dtype"], oracle.ref_setup)
        st = c["spec"].settings

        def fn(prob, settings, x0, Xref, Uref, state, cold, want):
            return oracle.solve_batch(prob, settings, x0, Xref, Uref, state=state, cold_start=cold, want_state=want,
                                      impl="reference")

        res, x0s = H.closed_loop(prob, st, c["inst"], c["steps"], c["reset_duals"], c["state"], fn)
        out = dict(nx=prob.nx, nu=prob.nu, N=prob.N, rho=prob.rho, dtype=np.dtype(prob.dtype).name, steps=c["steps"],
                   reset_duals=int(c["reset_duals"]), state_names=np.array(c["state"]))
        for f in PROBLEM_FIELDS:
            v = getattr(prob, f)
            if v is not None:
                out["prob_" + f] = v
        for n, _ in abi.Settings._fields_:
            out["set_" + n] = getattr(st, n)
        out["Xref"] = np.asarray(c["inst"]["Xref"], dtype=prob.dtype)
        if c["inst"].get("Uref") is not None:
            out["Uref"] = np.asarray(c["inst"]["Uref"], dtype=prob.dtype)
        out["x0_seq"] = np.stack(x0s)
        for k, r in enumerate(res):
            for key in H.OUT_KEYS + c["state"]:
                out[f"step{k}_{key}"] = r[key]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path)} bytes, iters step0 {res[0]['iter'].tolist()}")


if __name__ == "__main__":
    main()
