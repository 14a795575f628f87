"""Host-side (numpy) construction of a `tinympc_batch_t` (include/tinympc_b200.h).

Layout of every trajectory buffer: [B][N][nx] / [B][N-1][nu], C-contiguous = the reference's column-major
nx x N matrix (types.hpp:94-95) repeated B times.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .problem import MPCProblem


class HostBatch:
    """Owns the numpy buffers of one batched solve and the ctypes struct pointing at them."""

    def __init__(self, prob: MPCProblem, x0, Xref, Uref=None, state: dict | None = None, cold_start=True,
                 want_state=(), want_residuals=True, models=None):
        dt = prob.dtype
        nx, nu, N = prob.nx, prob.nu, prob.N
        self.prob = prob
        self.x0 = np.ascontiguousarray(x0, dtype=dt).reshape(-1, nx)
        B = self.x0.shape[0]
        self.B = B
        Xref = np.ascontiguousarray(Xref, dtype=dt)
        self.xref_per_instance = Xref.ndim == 3
        self.Xref = Xref.reshape((B, N, nx) if self.xref_per_instance else (N, nx))
        if Uref is None:
            self.Uref, self.uref_per_instance = None, False
        else:
            Uref = np.ascontiguousarray(Uref, dtype=dt)
            self.uref_per_instance = Uref.ndim == 3
            self.Uref = Uref.reshape((B, N - 1, nu) if self.uref_per_instance else (N - 1, nu))
        # state: in/out.  Arrays given in `state` are used in place (and updated); names in `want_state`
        # are allocated as zeros.
        self.state = {}
        for name in abi.STATE_FIELDS:
            shape = (B, N, nx) if abi.STATE_IS_X[name] else (B, N - 1, nu)
            if state is not None and name in state and state[name] is not None:
                a = state[name]
                if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.shape == shape):
                    a = np.ascontiguousarray(a, dtype=dt).reshape(shape).copy()
                self.state[name] = a
            elif name in want_state:
                self.state[name] = np.zeros(shape, dtype=dt)
        self.cold_start = bool(cold_start)
        self.models = None if models is None else np.ascontiguousarray(models, dtype=dt).reshape(B, -1)
        self.sol_x = np.zeros((B, N, nx), dtype=dt)
        self.sol_u = np.zeros((B, N - 1, nu), dtype=dt)
        self.iter = np.zeros(B, dtype=np.int32)
        self.solved = np.zeros(B, dtype=np.int32)
        self.residuals = np.zeros((B, 4), dtype=dt) if want_residuals else None

    def to_c(self) -> abi.Batch:
        b = abi.Batch()
        b.B = self.B
        b.x0 = self.x0.ctypes.data
        b.Xref = self.Xref.ctypes.data
        b.xref_per_instance = int(self.xref_per_instance)
        b.Uref = None if self.Uref is None else self.Uref.ctypes.data
        b.uref_per_instance = int(self.uref_per_instance)
        b.cold_start = int(self.cold_start)
        for name, a in self.state.items():
            setattr(b.state, name, a.ctypes.data)
        b.sol_x = self.sol_x.ctypes.data
        b.sol_u = self.sol_u.ctypes.data
        b.iter = self.iter.ctypes.data
        b.solved = self.solved.ctypes.data
        b.residuals = None if self.residuals is None else self.residuals.ctypes.data
        b.models = None if self.models is None else self.models.ctypes.data
        b._owner = self
        return b

    def result(self) -> dict:
        out = dict(sol_x=self.sol_x, sol_u=self.sol_u, iter=self.iter, solved=self.solved, residuals=self.residuals)
        out.update(self.state)
        return out
