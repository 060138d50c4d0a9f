"""Decomposed / sharded execution on the GPU: a slab-decomposed grid must reproduce the
undecomposed solve bit for bit (the per-cell arithmetic does not depend on the partition),
and a row-sharded ensemble must reproduce the unsharded one.  The multi-process tests need
>= 2 GPUs (gpurun --gpus 2) and are skipped on a single-GPU box."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ic(n: int, g: int) -> np.ndarray:
    x = (np.arange(n) + 0.5) / n
    return 0.5 + np.sin(2 * np.pi * x) + 0.3 * np.cos(6 * np.pi * x + 0.3)


def _reference_periodic(n: int, dt: float, nsteps: int, math: str) -> torch.Tensor:
    from pyshocks_b200.ensemble import EnsembleSolver

    g = 3
    s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g,
                       dx=3.0 / n, eps=1e-12, batch=1, math=math)
    u0 = np.zeros((1, n + 2 * g))
    u0[0, g : g + n] = _ic(n, g)
    s.solve_fixed_dt(torch.from_numpy(u0).cuda(), dt, nsteps)
    return s.u[0, g : g + n].clone()


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_local_slabs_equal_undecomposed(world: int, math: str) -> None:
    from pyshocks_b200.distributed import fill_halos_local, shard_rows
    from pyshocks_b200.ensemble import EnsembleSolver

    n, g, nsteps = 4099, 3, 12
    dt = 0.4 * (3.0 / n) / 1.8
    ref = _reference_periodic(n, dt, nsteps, math)
    ug = torch.from_numpy(_ic(n, g)).cuda()
    solvers = []
    for r in range(world):
        first, nl = shard_rows(n, r, world)
        s = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="none", n=nl, g=g,
                           dx=3.0 / n, eps=1e-12, batch=1, math=math)
        s.u[0, g : g + nl] = ug[first : first + nl]
        solvers.append(s)
    dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
    for _ in range(nsteps):
        fill_halos_local([s.u for s in solvers], g)
        for s in solvers:
            s.hp.stage(1, s.u, s.u, s.k1, dtt)
        fill_halos_local([s.k1 for s in solvers], g)
        for s in solvers:
            s.hp.stage(2, s.u, s.k1, s.k2, dtt)
        fill_halos_local([s.k2 for s in solvers], g)
        for s in solvers:
            s.hp.stage(3, s.u, s.k2, s.u, dtt)
    out = torch.cat([s.u[0, g : g + s.n] for s in solvers])
    assert torch.equal(out, ref)


@pytest.mark.parametrize("mode", ["serial", "overlap", "fused"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_peer_memory_slabs_equal_undecomposed(world: int, mode: str) -> None:
    """The peer-memory protocol with every slab held by this process: bit-identical to the
    periodic solve.  serial: psk_halo_wait -> stage -> psk_halo_push; overlap: slab edges on a
    second stream; fused: the exchange inside the stage kernel (psk_ssprk33_stage_p2p)."""
    from pyshocks_b200.distributed import PeerRing, PeerSlabSolver

    n, g, nsteps = (6144 if mode == "fused" else 6151), 3, 9
    dt = 0.4 * (3.0 / n) / 1.8
    ref = _reference_periodic(n, dt, nsteps, "fast")
    ug = torch.from_numpy(_ic(n, g)).cuda()
    slabs = [PeerSlabSolver(n_global=n, rank=r, world=world, dx=3.0 / n, edge=256, overlap=(mode == "overlap"),
                            fused=(mode == "fused"), timeout_s=5.0)
             for r in range(world)]
    try:
        assert all(s.split == (mode == "overlap") and s.fused == (mode == "fused") for s in slabs)
        for r, s in enumerate(slabs):
            s.attach(PeerRing.local([t.mem for t in slabs], r))
        for s in slabs:
            s.load_interior(ug[s.first : s.first + s.n_local])
        dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
        for _ in range(nsteps):
            for stage in (1, 2, 3):  # stage by stage: every wait depends on pushes enqueued before it
                for s in slabs:
                    s.run_stage(stage, dtt)
        for s in slabs:
            s.join()
            s.check()
        out = torch.cat([s.interior() for s in slabs])
        assert torch.equal(out, ref)
        assert all(s.exchanges == 3 * nsteps + 1 for s in slabs)
    finally:
        for s in slabs:
            s.ring = None
            s.solver = None
            s.mem.close()


@pytest.mark.parametrize("flux", ["rusanov", "eo"])
@pytest.mark.parametrize("world,n,nsteps", [(1, 6151, 5), (2, 6151, 8), (3, 4099, 7), (4, 500, 6), (2, 2 * 172 * 9 + 8, 6),
                                            (1, 172 * 4 + 3, 5)])
@pytest.mark.parametrize("fused_step", [False, True])
def test_whole_step_slabs_equal_undecomposed(world: int, n: int, nsteps: int, flux: str, fused_step: bool) -> None:
    """whole_step=True: ONE launch (psk_ssprk33_step on a slab with 9 ghost cells) and ONE exchange of 9
    cells per side per step, every slab held by this process: bit-identical to the periodic solve.
    fused_step: the exchange lives inside the kernel (psk_ssprk33_step_p2p), one launch per step in all."""
    from pyshocks_b200.distributed import PeerRing, PeerSlabSolver
    from pyshocks_b200.ensemble import EnsembleSolver

    g = 3
    dt = 0.4 * (3.0 / n) / 1.8
    single = EnsembleSolver(equation="burgers", flux=flux, rec="wenojs53", bc="periodic", n=n, g=g, dx=3.0 / n,
                            eps=1e-12, batch=1)
    u0 = np.zeros((1, n + 2 * g))
    u0[0, g : g + n] = _ic(n, g)
    single.solve_fixed_dt(torch.from_numpy(u0).cuda(), dt, nsteps)
    ref = single.u[0, g : g + n]
    ug = torch.from_numpy(_ic(n, g)).cuda()
    slabs = [PeerSlabSolver(n_global=n, rank=r, world=world, dx=3.0 / n, flux=flux, whole_step=True, timeout_s=5.0,
                            fused_step=None if fused_step else False)
             for r in range(world)]
    # psk_ssprk33_step_p2p needs slabs of at least two chunks of 172 cells
    covered = [s.n_local >= 344 for s in slabs]
    try:
        assert all(s.whole and s.g == 9 and not s.fused and not s.split for s in slabs)
        for r, s in enumerate(slabs):
            s.attach(PeerRing.local([t.mem for t in slabs], r))
        for s in slabs:
            s.load_interior(ug[s.first : s.first + s.n_local])
        dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
        for _ in range(nsteps):  # step by step: every wait depends on pushes enqueued before it
            for s in slabs:
                s.step(dtt)
        for s in slabs:
            s.join()
            s.check()
        out = torch.cat([s.interior() for s in slabs])
        assert torch.equal(out, ref)
        for s, cov in zip(slabs, covered):
            one_launch = fused_step and cov
            assert s.fused_step == one_launch
            assert s.exchanges == nsteps + 1 and s.launches == (1 if one_launch else 3) * nsteps + 1
    finally:
        for s in slabs:
            s.ring = None
            s.solver = None
            s.mem.close()


def test_whole_step_slab_adaptive() -> None:
    """timestepping.step's dt logic on a one-slab whole-step ring: the dt sequence and the bits of the
    single-array adaptive solve"""
    from pyshocks_b200.distributed import PeerRing, PeerSlabSolver
    from pyshocks_b200.ensemble import EnsembleSolver

    n, g = 1 << 13, 3
    ug = torch.from_numpy(_ic(n, g)).cuda()
    akw = dict(theta=0.8, tfinal=0.004, cfl_scale=0.5 * (3.0 / n))
    single = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g,
                            dx=3.0 / n, eps=1e-12, batch=1)
    u0 = torch.zeros((1, n + 2 * g), dtype=torch.float64, device="cuda")
    u0[0, g : g + n] = ug
    sres = single.solve_adaptive(u0, check_every=1, **akw)
    ps = PeerSlabSolver(n_global=n, rank=0, world=1, dx=3.0 / n, whole_step=True, timeout_s=5.0)
    try:
        ps.attach(PeerRing.local([ps.mem], 0))
        ps.load_interior(ug)
        pres = ps.solve_adaptive(**akw)
        assert pres.steps == sres.steps and float(ps.solver.t[0]) == float(single.t[0])
        assert torch.equal(ps.interior(), single.u[0, g : g + n])
    finally:
        ps.ring = None
        ps.solver = None
        ps.mem.close()


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_peer_memory_adaptive_single_slab(math: str) -> None:
    """timestepping.step's dt logic on a (one-slab) peer-memory ring: same dt sequence as the
    single-array adaptive solve; bit-identical state in STRICT mode (FAST: the slab runs the
    specialised kernel, the single-array solver the general one -- a few ulp per step)."""
    from pyshocks_b200.distributed import PeerRing, PeerSlabSolver
    from pyshocks_b200.ensemble import EnsembleSolver

    n, g = 1 << 14, 3
    ug = torch.from_numpy(_ic(n, g)).cuda()
    akw = dict(theta=0.8, tfinal=0.004, cfl_scale=0.5 * (3.0 / n))
    single = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g,
                            dx=3.0 / n, eps=1e-12, batch=1, math=math)
    u0 = torch.zeros((1, n + 2 * g), dtype=torch.float64, device="cuda")
    u0[0, g : g + n] = ug
    sres = single.solve_adaptive(u0, check_every=1, **akw)
    ps = PeerSlabSolver(n_global=n, rank=0, world=1, dx=3.0 / n, edge=512, math=math, timeout_s=5.0)
    try:
        ps.attach(PeerRing.local([ps.mem], 0))
        ps.load_interior(ug)
        pres = ps.solve_adaptive(**akw)
        assert ps.fused == (math == "fast") and ps.split == (math == "strict") and pres.steps == sres.steps
        assert float(ps.solver.t[0]) == float(single.t[0])
        diff = float((ps.interior() - single.u[0, g : g + n]).abs().max())
        assert diff == 0.0 if math == "strict" else diff <= 1e-13
    finally:
        ps.ring = None
        ps.solver = None
        ps.mem.close()


def test_fused_exchange_replayed_as_cuda_graph() -> None:
    """Fixed-dt loop of the fused exchange as a replayed CUDA graph of two steps (device-side epoch in
    two alternating slots) on a one-slab ring: bit-identical to the step-by-step solve, for an odd and
    an even number of steps and across two calls that reuse the captured graph."""
    from pyshocks_b200.distributed import PeerRing, PeerSlabSolver

    n, g = 6144, 3
    dt = 0.4 * (3.0 / n) / 1.8
    ug = torch.from_numpy(_ic(n, g)).cuda()
    ps = PeerSlabSolver(n_global=n, rank=0, world=1, dx=3.0 / n, timeout_s=5.0)
    try:
        assert ps.fused
        ps.attach(PeerRing.local([ps.mem], 0))
        ps.load_interior(ug)
        dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
        ps.solve_fixed_dt(dtt, 11, graph=True)
        ps.solve_fixed_dt(dtt, 8, graph=True)
        ps.check()
        assert ps._graph is not None and ps.exchanges == 3 * 19 + 1
        assert torch.equal(ps.interior(), _reference_periodic(n, dt, 19, "fast"))
    finally:
        ps.ring = None
        ps.solver = None
        ps._graph = None
        ps.mem.close()


def test_halo_wait_gives_up_instead_of_hanging() -> None:
    from pyshocks_b200 import _lib as L

    flags = torch.zeros(2, dtype=torch.int64, device="cuda")
    timed_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check("psk_halo_wait", L.lib().psk_halo_wait(flags.data_ptr(), flags.data_ptr() + 8, 1, int(0.05e9),
                                                   timed_out.data_ptr(), L.stream_ptr()))
    torch.cuda.synchronize()
    assert int(timed_out.item()) == 1
    flags.fill_(3)
    timed_out.zero_()
    L.check("psk_halo_wait", L.lib().psk_halo_wait(flags.data_ptr(), flags.data_ptr() + 8, 3, int(5e9),
                                                   timed_out.data_ptr(), L.stream_ptr()))
    torch.cuda.synchronize()
    assert int(timed_out.item()) == 0


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank: int, world: int, port: int, out: dict) -> None:
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pyshocks_b200.distributed import DistRing, PeerSlabSolver, ShardedEnsemble, SlabSolver, shard_rows
        from pyshocks_b200.ensemble import EnsembleSolver

        def mark(msg: str) -> None:
            print(f"[rank {rank}] {msg}", flush=True)

        mark("process group up")
        n, g, nsteps = 1 << 16, 3, 10
        dt = 0.4 * (3.0 / n) / 1.8
        ring = DistRing()
        ug = torch.from_numpy(_ic(n, g)).cuda()
        slab = SlabSolver(n_global=n, ring=ring, dx=3.0 / n)
        slab.load_interior(ug[slab.first : slab.first + slab.n_local])
        mark("slab built")
        slab.solve_fixed_dt(dt, nsteps)
        torch.cuda.synchronize()
        mark("slab fixed-dt done")
        ref = _reference_periodic(n, dt, nsteps, "fast")
        ok = torch.equal(slab.interior(), ref[slab.first : slab.first + slab.n_local])

        # peer-memory transport (CUDA IPC over NVLink), with and without the edge / interior overlap
        for mode in ("serial", "overlap", "fused"):
            ps = PeerSlabSolver(n_global=n, rank=rank, world=world, dx=3.0 / n, edge=1024, overlap=(mode == "overlap"),
                                fused=(mode == "fused"))
            ps.connect()
            ps.load_interior(ug[ps.first : ps.first + ps.n_local])
            ps.solve_fixed_dt(dt, nsteps)
            ps.check()
            ok = ok and torch.equal(ps.interior(), ref[ps.first : ps.first + ps.n_local])
            ok = ok and ps.split == (mode == "overlap") and ps.fused == (mode == "fused")
            if mode == "fused":  # ... and the same loop replayed as a CUDA graph, from the state reached so far
                ps.solve_fixed_dt(dt, 9, graph=True)
                ps.check()
                ref2 = _reference_periodic(n, dt, nsteps + 9, "fast")
                ok = ok and ps._graph is not None and torch.equal(ps.interior(), ref2[ps.first : ps.first + ps.n_local])
                ps._graph = None
            ps.close()
            mark(f"peer slabs {mode} ok={ok}")
        # whole step per launch on slabs with 9 ghost cells: wait / step / push, and the exchange inside the kernel
        for fused_step in (False, None):
            ps = PeerSlabSolver(n_global=n, rank=rank, world=world, dx=3.0 / n, whole_step=True, fused_step=fused_step)
            ps.connect()
            ps.load_interior(ug[ps.first : ps.first + ps.n_local])
            ps.solve_fixed_dt(dt, nsteps)
            ps.check()
            ok = ok and torch.equal(ps.interior(), ref[ps.first : ps.first + ps.n_local])
            ok = ok and ps.launches == (1 if ps.fused_step else 3) * nsteps + 1
            mark(f"whole-step slabs fused_step={ps.fused_step} launches={ps.launches} ok={ok}")
            ps.close()
        # discrete adjoint on the slabs against the gradient of the undecomposed periodic solve
        from pyshocks_b200.distributed import PeerSlabAdjoint
        from pyshocks_b200.ensemble import AdjointEnsemble

        na, ka = 1 << 14, 8
        dta = 0.4 * (3.0 / na) / 1.8
        uga = torch.from_numpy(_ic(na, g)).cuda()
        solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=na, g=g, dx=3.0 / na,
                                eps=1e-12, batch=1)
        u0a = torch.zeros((1, na + 2 * g), dtype=torch.float64, device="cuda")
        u0a[0, g : g + na] = uga
        _, g_ref = AdjointEnsemble(solver, nsteps=ka, dt=dta, segment=1).gradient_half_l2(u0a)
        sa = PeerSlabAdjoint(n_global=na, rank=rank, world=world, dx=3.0 / na, nsteps=ka)
        sa.connect()
        _, grad = sa.gradient_half_l2(uga[sa.first : sa.first + sa.n_local], dta)
        want = g_ref[0, g + sa.first : g + sa.first + sa.n_local]
        err = float((grad - want).abs().max() / g_ref.abs().max())
        ok = ok and err < 1e-12
        mark(f"slab adjoint: max rel err {err:.3e} ok={ok}")
        sa.close()
        # adaptive dt: every rank takes the same dt sequence as the single-GPU adaptive solve.  The
        # single-array solver runs the general kernel (row mask), the slabs the specialised one: two FAST
        # implementations agree to a few ulp per step, the STRICT ones bit for bit; the two slab
        # transports run the same kernels on the same cells and must agree bit for bit in both modes.
        akw = dict(theta=0.8, tfinal=0.002, cfl_scale=0.5 * (3.0 / n))
        for math in ("strict", "fast"):
            pa = PeerSlabSolver(n_global=n, rank=rank, world=world, dx=3.0 / n, edge=1024, math=math)
            pa.connect()
            pa.load_interior(ug[pa.first : pa.first + pa.n_local])
            pres = pa.solve_adaptive(**akw)
            slab2 = SlabSolver(n_global=n, ring=ring, dx=3.0 / n, math=math)
            slab2.load_interior(ug[slab2.first : slab2.first + slab2.n_local])
            res = slab2.solve_adaptive(**akw)
            single = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g,
                                    dx=3.0 / n, eps=1e-12, batch=1, math=math)
            u0 = torch.zeros((1, n + 2 * g), dtype=torch.float64, device="cuda")
            u0[0, g : g + n] = ug
            sres = single.solve_adaptive(u0, check_every=1, **akw)
            mine = single.u[0, g + pa.first : g + pa.first + pa.n_local]
            diff = float((pa.interior() - mine).abs().max())
            mark(f"adaptive {math}: steps peer {pres.steps} / nccl {res.steps} / single {sres.steps}; max diff {diff:.3e}")
            ok = ok and res.steps == sres.steps and pres.steps == sres.steps
            ok = ok and torch.equal(pa.interior(), slab2.interior())
            ok = ok and (diff == 0.0 if math == "strict" else diff <= 1e-13)
            pa.close()

        # row-sharded ensemble == unsharded ensemble
        B, nn = 37, 512
        rng = np.random.default_rng(3)
        U = torch.from_numpy(0.5 + 0.5 * rng.standard_normal((B, nn + 2 * g))).cuda()
        kw = dict(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=nn, g=g, dx=3.0 / nn, eps=1e-12)
        sh = ShardedEnsemble(batch=B, **kw)
        sh.solver.solve_fixed_dt(sh.local_rows(U), 1e-4, 5)
        full = sh.gather(dst=0)
        mark("gather done")
        if rank == 0:
            one = EnsembleSolver(batch=B, **kw)
            one.solve_fixed_dt(U, 1e-4, 5)
            ok = ok and torch.equal(full[:, g : g + nn], one.u[:, g : g + nn])
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out["ok"] = int(flag) == 1
            out["exchanges"] = slab.exchanges
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_multi_gpu_slabs_and_sharded_ensemble() -> None:
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out.get("ok") is True
    assert out["exchanges"] == 30  # 3 halo exchanges per step


@pytest.mark.parametrize("world,n,nsteps", [(1, 1024, 6), (2, 2048, 9), (3, 1536, 7), (4, 128, 5)])
def test_slab_adjoint_equals_the_periodic_gradient(world: int, n: int, nsteps: int) -> None:
    """PeerSlabAdjoint on an in-process ring: forward sweep onto per-slab tapes (16 ghost cells filled by the
    neighbours' pushes) and the gathered reverse sweep, against AdjointEnsemble on the undecomposed periodic row.
    Forward states are bit-identical; the gradient agrees to round-off (the windows of the reverse kernel, hence
    the order in which a first difference collects its cotangents, start elsewhere on a slab)."""
    from pyshocks_b200.distributed import PeerSlabAdjoint
    from pyshocks_b200.ensemble import AdjointEnsemble, EnsembleSolver

    g = 3
    dx = 3.0 / n
    dt = 0.4 * dx / 1.8
    ug = torch.from_numpy(_ic(n, g)).cuda()
    solver = EnsembleSolver(equation="burgers", flux="rusanov", rec="wenojs53", bc="periodic", n=n, g=g, dx=dx,
                            eps=1e-12, batch=1)
    u0 = torch.zeros((1, n + 2 * g), dtype=torch.float64, device="cuda")
    u0[0, g : g + n] = ug
    ref = AdjointEnsemble(solver, nsteps=nsteps, dt=dt, segment=1)
    J_ref, g_ref = ref.gradient_half_l2(u0)
    uT_ref = ref.chk[-1][0, g : g + n]
    slabs = [PeerSlabAdjoint(n_global=n, rank=r, world=world, dx=dx, nsteps=nsteps, timeout_s=5.0) for r in range(world)]
    try:
        for s in slabs:
            s.attach_local(slabs)
        dtt = torch.full((1,), dt, dtype=torch.float64, device="cuda")
        for s in slabs:
            s.forward_begin(ug[s.first : s.first + s.n_local])
        for m in range(nsteps):  # step by step: every wait depends on pushes enqueued before it
            for s in slabs:
                s.forward_step(m, dtt)
        uT = torch.cat([s.interior(nsteps) for s in slabs])
        assert torch.equal(uT, uT_ref)
        for s in slabs:
            s.backward_begin(s.interior(nsteps))
        for m in range(nsteps - 1, -1, -1):
            for s in slabs:
                s.backward_step(m, dtt)
        for s in slabs:
            s.check()
        grad = torch.cat([s.interior(s._cur) for s in slabs])
        want = g_ref[0, g : g + n]
        assert float((grad - want).abs().max()) < 1e-12 * float(want.abs().max())
        assert all(s.launches == 3 * 2 * nsteps + 2 for s in slabs)
    finally:
        for s in slabs:
            s.close()
