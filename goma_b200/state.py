"""Synthetic state vectors of SURVEY.md §8d, shared by the parity tests, the golden-fixture generator and bench.py.

A smooth analytic field plus seeded noise, so that every Jacobian term is exercised:
u = (sin pi x cos pi y, -cos pi x sin pi y[, 0.3 sin pi z]) + 0.05 N(0,1); P ~ N(0,1); T = x + 0.1 N; Y_w = 0.5 + 0.05 N;
d = 0.02 h N; ``numpy.random.default_rng(seed)``; the transient variant adds x_old, xdot_old and the xdot of
``rf_solve.c:2848``.
"""
from __future__ import annotations

import numpy as np

SEED = 20261017


def make_state(problem, seed=SEED, transient=False, delta_t=0.01, theta=0.0):
    """Returns {"x": ...} (+ x_old, x_older, xdot_old, xdot when ``transient``)."""
    rng = np.random.default_rng(seed)
    m = problem.mesh
    first, node_kind, kinds = problem.unknown_map()
    n = int(first[-1])
    X = m.coords

    def fill():
        x = np.zeros(n)
        for kind_id, slots in enumerate(kinds):
            nodes = np.nonzero(node_kind == kind_id)[0]
            for off, name in enumerate(slots):
                idx = first[nodes] + off
                cx, cy = X[0][nodes], X[1][nodes]
                noise = rng.normal(size=len(nodes))
                if name == "U":
                    v = np.sin(np.pi * cx) * np.cos(np.pi * cy) + 0.05 * noise
                elif name == "V":
                    v = -np.cos(np.pi * cx) * np.sin(np.pi * cy) + 0.05 * noise
                elif name == "W":
                    v = 0.3 * np.sin(np.pi * X[2][nodes]) + 0.05 * noise
                elif name == "T":
                    v = cx + 0.1 * noise
                elif name.startswith("Y"):
                    v = 0.5 + 0.05 * noise
                elif name.startswith("D"):
                    h = 1.0 / max(max(m.lattice), 1)
                    v = 0.02 * h * noise
                else:  # pressure dofs
                    v = noise
                x[idx] = v
        return x

    st = {"x": fill()}
    if transient:
        x_old = st["x"] + 0.03 * rng.normal(size=n)
        xdot_old = rng.normal(size=n) * 0.1
        # rf_solve.c:2848: xdot = (1+2 theta)/dt (x - x_old) - 2 theta xdot_old
        st.update(x_old=x_old, x_older=x_old.copy(), xdot_old=xdot_old,
                  xdot=(1 + 2 * theta) / delta_t * (st["x"] - x_old) - 2 * theta * xdot_old)
    return st
