#!/usr/bin/env python
"""Record golden vectors of the hot path from the reference's own Python code.

Run in the development container only (``/root/reference`` does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference is imported from ``/root/reference/src`` with the NumPy-backed
``jax`` stand-in of ``tests/golden/_jaxshim`` on ``sys.path`` (JAX itself is not
installable here) and ``importlib.metadata.version("pyshocks")`` stubbed
(``pyshocks/__init__.py:58`` needs an installed distribution).  The reference's
code is executed unmodified; only the array backend differs from real JAX.

Outputs (all small, committed):

* ``weno.npz``     weno_smoothness / weno_interp / weno_js_weights / reconstruct
* ``rhs.npz``      apply_boundary, numerical_flux, apply_operator, predict_timestep
                   for every case of ``cases.rhs_cases()``
* ``advance.npz``  one SSPRK33 ``advance`` per case (subset)
* ``steppers.npz`` one ``advance`` of ForwardEuler / RK44 / CKRK45 (timestepping.py:289-405)
                   for the cases of ``cases.stepper_cases()``
* ``solve_c1.npz`` BASELINE config 1 (examples/burgers.py, N=256, t=1) for the
                   rusanov and lf schemes: dt history and final state
* ``solve_c2.npz`` config-2 forward at reduced N (advection, Dirichlet, theta=.75)
* ``solve_c2_4096.npz`` config-2 forward at its full size N = 4096 (2731 steps): dt history, three
                   intermediate states and the final state
* ``adjoint.npz``  ``adjoint_step`` sweeps (burgers-adjoint / advection-adjoint
                   driver set-ups at small N), every intermediate ``p``
"""

from __future__ import annotations

import importlib.metadata as md
import pathlib
import sys
from functools import partial

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "_jaxshim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(HERE))

_version = md.version
md.version = lambda name: "0.0.0+shim" if name == "pyshocks" else _version(name)

import jax.numpy as jnp  # noqa: E402  (the shim)

import cases as C  # noqa: E402
import pyshocks  # noqa: E402
from pyshocks import (  # noqa: E402
    advection,
    apply_boundary,
    apply_operator,
    bind,
    burgers,
    cell_average,
    continuity,
    funcs,
    make_leggauss_quadrature,
    make_uniform_cell_grid,
    numerical_flux,
    predict_timestep,
    timestepping,
)
from pyshocks.checkpointing import InMemoryCheckpoint  # noqa: E402
from pyshocks.reconstruction import make_reconstruction_from_name, reconstruct  # noqa: E402
from pyshocks.scalar import (  # noqa: E402
    PeriodicBoundary,
    make_dirichlet_boundary,
    make_neumann_boundary,
)
from pyshocks.schemes import BoundaryType  # noqa: E402

A = np.asarray


def build(case: C.Case):
    rec = make_reconstruction_from_name(case.rec)
    grid = make_uniform_cell_grid(a=case.a, b=case.b, n=case.n, nghosts=case.g)
    if case.equation == "burgers":
        scheme = burgers.make_scheme_from_name(case.flux, rec=rec, alpha=case.alpha)
    elif case.equation == "advection":
        scheme = advection.make_scheme_from_name(
            case.flux, rec=rec, velocity=jnp.array(C.velocity_for(case))
        )
    else:
        scheme = continuity.make_scheme_from_name(
            case.flux, rec=rec, velocity=jnp.array(C.velocity_for(case))
        )
    if case.bc == "periodic":
        bc = PeriodicBoundary()
    else:
        bc = make_dirichlet_boundary(
            ga=lambda t, x: jnp.array(C.dirichlet_values(case, float(t), A(x)))
        )
    scheme = bind(scheme, grid, bc)
    return scheme, grid, bc


def golden_weno() -> dict[str, np.ndarray]:
    from pyshocks.weno import weno_interp, weno_js_weights, weno_smoothness

    out: dict[str, np.ndarray] = {}
    for name in ("wenojs32", "wenojs53"):
        rec = make_reconstruction_from_name(name)
        g = rec.stencil_width
        n = 64 + 2 * g
        theta = np.linspace(0.0, 2.0 * np.pi, n)
        for label, u in (
            ("sine", np.sin(theta)),
            ("step", (theta < np.pi).astype(np.float64)),
            ("rough", np.sin(3 * theta) + (theta > 2.0) * 0.7 + 1e-3 * np.cos(40 * theta)),
        ):
            uj = jnp.array(u)
            k = f"{name}_{label}"
            out[f"{k}_u"] = u
            out[f"{k}_beta"] = A(weno_smoothness(rec.s, uj))
            out[f"{k}_uhat"] = A(weno_interp(rec.s, uj))
            out[f"{k}_omega"] = A(weno_js_weights(rec.s, uj, eps=rec.eps))
            grid = make_uniform_cell_grid(a=0.0, b=1.0, n=64, nghosts=g)
            ul, ur = reconstruct(rec, grid, BoundaryType.Dirichlet, uj, uj, uj)
            out[f"{k}_ul"] = A(ul)
            out[f"{k}_ur"] = A(ur)
    # ESWENO32 weights (weno.py:284-296) and reconstruction (reconstruction.py:413-439)
    from pyshocks.weno import es_weno_weights

    rec = make_reconstruction_from_name("esweno32")
    g = rec.stencil_width
    n = 64 + 2 * g
    theta = np.linspace(0.0, 2.0 * np.pi, n)
    for label, u in (
        ("sine", np.sin(theta)),
        ("step", (theta < np.pi).astype(np.float64)),
        ("rough", np.sin(3 * theta) + (theta > 2.0) * 0.7 + 1e-3 * np.cos(40 * theta)),
    ):
        uj = jnp.array(u)
        k = f"esweno32_{label}"
        out[f"{k}_u"] = u
        out[f"{k}_omega"] = A(es_weno_weights(rec.s, uj, eps=rec.eps))
        grid = make_uniform_cell_grid(a=0.0, b=1.0, n=64, nghosts=g)
        ul, ur = reconstruct(rec, grid, BoundaryType.Dirichlet, uj, uj, uj)
        out[f"{k}_ul"] = A(ul)
        out[f"{k}_ur"] = A(ur)
    return out


def golden_rhs() -> tuple[dict[str, np.ndarray], dict[str, np.ndarray]]:
    rhs: dict[str, np.ndarray] = {}
    adv: dict[str, np.ndarray] = {}
    for case in C.rhs_cases():
        scheme, grid, bc = build(case)
        u = jnp.array(C.state_for(case))
        k = case.key
        rhs[f"{k}_u"] = A(u)
        w = apply_boundary(bc, grid, case.t, u)
        rhs[f"{k}_w"] = A(w)
        if case.equation != "advection" or True:
            rhs[f"{k}_f"] = A(numerical_flux(scheme, grid, bc, case.t, w))
        rhs[f"{k}_L"] = A(apply_operator(scheme, grid, bc, case.t, u))
        rhs[f"{k}_dt"] = A(predict_timestep(scheme, grid, bc, case.t, u))

        # one SSPRK33 step with a CFL-like dt (subset: rough states only)
        if case.rec == "esweno32":
            rhs[f"{k}_eps"] = np.float64(scheme.rec.eps)
            rhs[f"{k}_delta"] = np.float64(scheme.rec.delta)
        if case.state == "rough" or case.rec in ("wenojs53", "esweno32"):
            dt = 0.3 * float(A(predict_timestep(scheme, grid, bc, case.t, u)))
            stepper = timestepping.SSPRK33(
                predict_timestep=lambda t_, u_: dt,
                source=partial(apply_operator, scheme, grid, bc),
                checkpoint=None,
            )
            adv[f"{k}_u"] = A(u)
            adv[f"{k}_dt"] = np.float64(dt)
            adv[f"{k}_out"] = A(timestepping.advance(stepper, dt, case.t, u))
    return rhs, adv


def golden_steppers() -> dict[str, np.ndarray]:
    """timestepping.py:289-301 (ForwardEuler), :328-343 (RK44), :352-405 (CKRK45): one ``advance`` each"""
    out: dict[str, np.ndarray] = {}
    for case in C.stepper_cases():
        scheme, grid, bc = build(case)
        u = jnp.array(C.state_for(case))
        dt = 0.3 * float(A(predict_timestep(scheme, grid, bc, case.t, u)))
        out[f"{case.key}_u"] = A(u)
        out[f"{case.key}_dt"] = np.float64(dt)
        for name in C.STEPPERS:
            stepper = getattr(timestepping, name)(
                predict_timestep=lambda t_, u_: dt,
                source=partial(apply_operator, scheme, grid, bc),
                checkpoint=None,
            )
            out[f"{case.key}_{name}"] = A(timestepping.advance(stepper, dt, case.t, u))
    return out


def golden_solve_c1() -> dict[str, np.ndarray]:
    """examples/burgers.py:42-60,120-179 with -s rusanov|lf -r wenojs53 -n 256, tfinal=1."""
    out: dict[str, np.ndarray] = {}
    for sname in ("rusanov", "lf"):
        rec = make_reconstruction_from_name("wenojs53")
        scheme = burgers.make_scheme_from_name(sname, rec=rec, alpha=1.0)
        order = int(max(scheme.order, 1.0)) + 1
        grid = make_uniform_cell_grid(a=-1.5, b=1.5, n=256, nghosts=scheme.stencil_width)
        quad = make_leggauss_quadrature(grid, order=order)
        u0 = cell_average(quad, lambda x: funcs.burgers_tophat(grid, 0.0, x))
        bc = PeriodicBoundary()
        scheme = bind(scheme, grid, bc)
        theta = 1.0
        method = timestepping.SSPRK33(
            predict_timestep=lambda t_, u_: theta * predict_timestep(scheme, grid, bc, t_, u_),
            source=lambda t_, u_: apply_operator(scheme, grid, bc, t_, u_),
            checkpoint=None,
        )
        dts, ts, snaps = [], [], {}
        for event in timestepping.step(method, u0, tfinal=1.0):
            dts.append(float(event.dt))
            ts.append(float(event.t))
            if event.iteration in (1, 10, 50, 100):
                snaps[event.iteration] = A(event.u)
        out[f"{sname}_u0"] = A(u0)
        out[f"{sname}_dt"] = np.array(dts)
        out[f"{sname}_t"] = np.array(ts)
        out[f"{sname}_uf"] = A(event.u)
        for m, v in snaps.items():
            out[f"{sname}_u{m:03d}"] = v
    return out


def _advection_driver(n: int, bctype: str = "dirichlet", rec_name: str = "wenojs53"):
    """drivers/advection-adjoint.py:219-294."""
    rec = make_reconstruction_from_name(rec_name)
    scheme = advection.make_scheme_from_name("godunov", rec=rec, velocity=None)
    grid = make_uniform_cell_grid(a=-1.0, b=1.0, n=n, nghosts=scheme.stencil_width)
    order = int(max(scheme.order, 1.0)) + 1
    quad = make_leggauss_quadrature(grid, order=order)
    from dataclasses import replace

    velocity = cell_average(quad, partial(funcs.ic_constant, grid, c=1.0))
    scheme = replace(scheme, velocity=velocity)
    func_ic = partial(funcs.ic_sine, grid, k=1)
    u0 = cell_average(quad, func_ic)
    if bctype == "periodic":
        bc = PeriodicBoundary()
    else:
        bc = make_dirichlet_boundary(lambda t, x: func_ic(x - 1.0 * t))
    theta = 0.75
    stepper = timestepping.SSPRK33(
        predict_timestep=lambda t_, u_: theta * predict_timestep(scheme, grid, bc, t_, u_),
        source=lambda t_, u_: apply_operator(scheme, grid, bc, t_, u_),
        checkpoint=InMemoryCheckpoint(basename="Iteration"),
    )
    return scheme, grid, bc, stepper, u0, velocity


def golden_solve_c2() -> dict[str, np.ndarray]:
    out: dict[str, np.ndarray] = {}
    scheme, grid, bc, stepper, u0, velocity = _advection_driver(128)
    dts = []
    for event in timestepping.step(stepper, u0, tfinal=1.0):
        dts.append(float(event.dt))
    out["u0"] = A(u0)
    out["velocity"] = A(velocity)
    out["dt"] = np.array(dts)
    out["uf"] = A(event.u)
    return out


def golden_solve_c2_full() -> dict[str, np.ndarray]:
    """BASELINE config 2 at its full size: drivers/advection-adjoint.py:219-276 with -s godunov -r wenojs53
    -n 4096 (Dirichlet exact-solution boundary, theta = 0.75, t = 1): 2731 steps of the forward solve."""
    out: dict[str, np.ndarray] = {}
    scheme, grid, bc, stepper, u0, velocity = _advection_driver(4096)
    stepper = timestepping.SSPRK33(predict_timestep=stepper.predict_timestep, source=stepper.source, checkpoint=None)
    dts = []
    for event in timestepping.step(stepper, u0, tfinal=1.0):
        dts.append(float(event.dt))
        if event.iteration in (1, 100, 1000):
            out[f"u{event.iteration:04d}"] = A(event.u)
    out["u0"] = A(u0)
    out["dt"] = np.array(dts)
    out["uf"] = A(event.u)
    return out


def _record_adjoint(out, key, stepper, grid, u0, tfinal, p_boundary):
    for event in timestepping.step(stepper, u0, tfinal=tfinal):
        pass
    uf, maxit = event.u, event.iteration
    chk = stepper.checkpoint
    out[f"{key}_u0"] = A(u0)
    out[f"{key}_maxit"] = np.int64(maxit)
    out[f"{key}_chk_t"] = np.array([float(chk.storage[("Iteration", m)]["t"]) for m in range(maxit + 1)])
    out[f"{key}_chk_u"] = np.stack([A(chk.storage[("Iteration", m)]["u"]) for m in range(maxit + 1)])
    ps, dts = [], []
    for event in timestepping.adjoint_step(
        stepper, uf, maxit=maxit, apply_boundary=lambda t, u, p: apply_boundary(p_boundary, grid, t, p)
    ):
        ps.append(A(event.p))
        dts.append(float(event.dt))
    out[f"{key}_p"] = np.stack(ps)  # p[0] = p(T) after BC, p[-1] = p(0)
    out[f"{key}_adj_dt"] = np.array(dts)


def golden_adjoint() -> dict[str, np.ndarray]:
    out: dict[str, np.ndarray] = {}

    # drivers/burgers-adjoint.py:68-97,205-212: Dirichlet tophat forward BC,
    # homogeneous Neumann BC on p, p(T) = u(T)
    for sname, rec_name, alpha in (
        ("rusanov", "wenojs53", 1.0),
        ("lf", "wenojs53", 1.0),
        ("godunov", "wenojs53", 1.0),
        ("eo", "wenojs53", 1.0),
        ("rusanov", "wenojs32", 1.0),
        ("lf", "constant", 0.995),
    ):
        rec = make_reconstruction_from_name(rec_name)
        scheme = burgers.make_scheme_from_name(sname, rec=rec, alpha=alpha)
        grid = make_uniform_cell_grid(a=-1.5, b=1.5, n=48, nghosts=scheme.stencil_width)
        order = int(max(scheme.order, 1)) + 1
        quad = make_leggauss_quadrature(grid, order=order)
        u0 = cell_average(quad, lambda x: funcs.burgers_tophat(grid, 0.0, x))
        bc = make_dirichlet_boundary(ga=lambda t, x: funcs.burgers_tophat(grid, t, x))
        stepper = timestepping.SSPRK33(
            predict_timestep=lambda t_, u_: 1.0 * predict_timestep(scheme, grid, bc, t_, u_),
            source=lambda t_, u_: apply_operator(scheme, grid, bc, t_, u_),
            checkpoint=InMemoryCheckpoint(basename="Iteration"),
        )
        scheme = bind(scheme, grid, bc)
        pbc = make_neumann_boundary(lambda t: jnp.array(0.0))
        key = f"burgers_{sname}_{rec_name}" + ("" if alpha == 1.0 else f"_alpha{alpha:g}")
        _record_adjoint(out, key, stepper, grid, u0, 0.4, pbc)

    # same scheme, periodic forward BC and periodic BC on p
    rec = make_reconstruction_from_name("wenojs53")
    scheme = burgers.make_scheme_from_name("rusanov", rec=rec, alpha=1.0)
    grid = make_uniform_cell_grid(a=-1.5, b=1.5, n=48, nghosts=3)
    quad = make_leggauss_quadrature(grid, order=4)
    u0 = cell_average(quad, lambda x: 0.5 - funcs.ic_sine(grid, x))
    bc = PeriodicBoundary()
    stepper = timestepping.SSPRK33(
        predict_timestep=lambda t_, u_: predict_timestep(scheme, grid, bc, t_, u_),
        source=lambda t_, u_: apply_operator(scheme, grid, bc, t_, u_),
        checkpoint=InMemoryCheckpoint(basename="Iteration"),
    )
    _record_adjoint(out, "burgers_rusanov_wenojs53_periodic", stepper, grid, u0, 0.4, bc)

    # drivers/advection-adjoint.py: Dirichlet exact-solution BC forward,
    # Dirichlet zeros on p
    for bctype in ("dirichlet", "periodic"):
        scheme, grid, bc, stepper, u0, _ = _advection_driver(48, bctype=bctype)
        pbc = (
            make_dirichlet_boundary(lambda t, x: jnp.zeros_like(x))
            if bctype == "dirichlet"
            else PeriodicBoundary()
        )
        _record_adjoint(out, f"advection_godunov_wenojs53_{bctype}", stepper, grid, u0, 0.5, pbc)
    return out


def main() -> None:
    print("reference:", pyshocks.__file__)
    if sys.argv[1:] == ["steppers"]:  # this file alone (the others are untouched by it)
        np.savez_compressed(HERE / "steppers.npz", **golden_steppers())
        return
    if sys.argv[1:] == ["c2full"]:
        np.savez_compressed(HERE / "solve_c2_4096.npz", **golden_solve_c2_full())
        return
    np.savez_compressed(HERE / "weno.npz", **golden_weno())
    rhs, adv = golden_rhs()
    np.savez_compressed(HERE / "rhs.npz", **rhs)
    np.savez_compressed(HERE / "advance.npz", **adv)
    np.savez_compressed(HERE / "steppers.npz", **golden_steppers())
    np.savez_compressed(HERE / "solve_c1.npz", **golden_solve_c1())
    np.savez_compressed(HERE / "solve_c2.npz", **golden_solve_c2())
    np.savez_compressed(HERE / "solve_c2_4096.npz", **golden_solve_c2_full())
    np.savez_compressed(HERE / "adjoint.npz", **golden_adjoint())
    for f in sorted(HERE.glob("*.npz")):
        print(f"{f.name}: {f.stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
