"""The drop-in boundary driven from plain C (no Python in the process that solves): tests/capi/c_abi_batch.c dlopens
libtinympc_b200.so, resolves the header's entry points by name and solves a batch; its dump is compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import oracle
from tinympc_b200 import _lib, workloads as wl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_host_solves_a_batch_bit_identical_to_oracle(tmp_path):
    exe = str(tmp_path / "c_abi_batch")
    src = os.path.join(ROOT, "tests", "capi", "c_abi_batch.c")
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-ldl"], check=True)
    B, nx, nu, N = 1000, 4, 1, 10
    out = str(tmp_path / "dump.bin")
    r = subprocess.run([exe, _lib.LIB_PATH, str(B), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "c_abi_batch: B=1000" in r.stdout
    raw = np.fromfile(out, dtype=np.uint8)
    o = 0

    def take(dt, n):
        nonlocal o
        a = raw[o:o + n * np.dtype(dt).itemsize].view(dt)
        o += n * np.dtype(dt).itemsize
        return a
    x0 = take(np.float64, B * nx).reshape(B, nx)
    got = dict(sol_x=take(np.float64, B * N * nx).reshape(B, N, nx), sol_u=take(np.float64, B * (N - 1) * nu).reshape(B, N - 1, nu),
               iter=take(np.int32, B), solved=take(np.int32, B), residuals=take(np.float64, B * 4).reshape(B, 4))
    assert o == raw.size
    spec = wl.cartpole()
    spec.constraints = dict(x_min=np.full(4, -5.0), x_max=np.full(4, 5.0), u_min=np.full(1, -3.0), u_max=np.full(1, 3.0))
    from tinympc_b200.solver import setup_problem
    prob = setup_problem(spec, np.float64)  # same host precompute the C program called
    ref = oracle.solve_batch(prob, spec.settings, x0, np.tile([1.0, 0, 0, 0], (N, 1)), None, cold_start=True, impl="port", nthreads=4)
    for k in H.OUT_KEYS:
        assert H.bits_equal(got[k], ref[k]), k
    assert got["solved"].all() and got["iter"].min() >= 1
