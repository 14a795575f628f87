"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: contiguous sharding covers the batch exactly,
results are shard-invariant (each rank solves its shard — with the oracle standing in for the GPU — and the
gathered result equals the single-process solve bit for bit), and the post-solve statistics reduction
(sum of instances / solved / iterations, max of residuals / ms) is correct."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from oracle import oracle
    from tinympc_b200 import workloads as wl
    from tinympc_b200.parallel import reduce_stats, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = wl.quadrotor(N=10)
    prob = oracle.port_setup(spec.nx, spec.nu, spec.N, spec.rho, spec.A, spec.B, spec.f, spec.Qdiag, spec.Rdiag,
                             dtype=np.float32, **spec.constraints)
    inst = wl.tracking_instances(B, N=10, seed=3, dtype=np.float32)
    lo, hi = shard_bounds(B, rank, world)
    r = oracle.solve_batch(prob, spec.settings, inst["x0"][lo:hi], inst["Xref"][lo:hi], None, cold_start=True, impl="port")
    stats = dict(instances=hi - lo, solved=int(r["solved"].sum()), iters=int(r["iter"].sum()),
                 res_max=r["residuals"].max(axis=0).astype(np.float64).tolist(), ms=10.0 + rank)
    red = reduce_stats(stats)
    q.put((rank, lo, hi, r["sol_u"], r["iter"], red))
    dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    from tinympc_b200.parallel import shard_bounds

    for B in (1, 7, 64, 65536, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= (B + world - 1) // world


def test_world2_gloo_shard_invariance_and_stats():
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    from oracle import oracle
    from tinympc_b200 import workloads as wl

    B, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    spec = wl.quadrotor(N=10)
    prob = oracle.port_setup(spec.nx, spec.nu, spec.N, spec.rho, spec.A, spec.B, spec.f, spec.Qdiag, spec.Rdiag,
                             dtype=np.float32, **spec.constraints)
    inst = wl.tracking_instances(B, N=10, seed=3, dtype=np.float32)
    full = oracle.solve_batch(prob, spec.settings, inst["x0"], inst["Xref"], None, cold_start=True, impl="port")
    sol_u = np.concatenate([g[3] for g in got])
    iters = np.concatenate([g[4] for g in got])
    assert np.array_equal(sol_u.view(np.uint8), full["sol_u"].view(np.uint8))
    assert np.array_equal(iters, full["iter"])
    for g in got:
        red = g[5]
        assert red["instances"] == B and red["iters"] == int(full["iter"].sum()) and red["solved"] == int(full["solved"].sum())
        assert red["ms"] == 11.0
        assert np.allclose(red["res_max"], full["residuals"].max(axis=0).astype(np.float64))
