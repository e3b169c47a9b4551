"""A minimal host Newton loop around one matrix_fill implementation -- TEST INFRASTRUCTURE.

Mirrors the part of solve_nonlinear_problem (``src/mm_sol_nonlinear.c:1104-1600``) that surrounds the
hot path: zero, fill (residual + Jacobian), linear solve J dx = R, x -= dx (:2074, Newton correction
factor 1), until the L2 norm of the residual is below the Normalized Residual Tolerance.  The linear
solve is scipy's sparse LU (the reference's is Amesos/Aztec -- out of scope); what the
tests compare is the sequence of assembled systems through its effect: iteration count and solution.
"""
import numpy as np
import scipy.sparse.linalg as spla

from goma_b200.matrix_fill import msr_to_csr


def newton(fill, ija, x0, tol=1e-10, max_it=15):
    """``fill(x) -> (err, a, resid)`` in the MSR layout of ``ija``.  Returns (x, iterations, residual norms)."""
    x = x0.copy()
    n = len(x)
    norms = []
    for it in range(max_it):
        err, a, r = fill(x)
        assert err == 0
        norms.append(float(np.linalg.norm(r)))
        if norms[-1] < tol:
            return x, it, norms
        A = msr_to_csr(ija, a, n)
        x = x - spla.spsolve(A.tocsc(), r)
    return x, max_it, norms


def channel_problem(n=(40, 10), length=4.0):
    """BASELINE.json configs[0] / SURVEY.md §8d C1: 2-D steady Newtonian channel, Q2/P1 on a 40x10 quad9 mesh,
    mu = rho = 1; inflow U=1,V=0, walls U=V=0, outflow V=0 (constant Dirichlet cards only)."""
    from goma_b200.mesh import box_mesh
    from goma_b200.problem import Dirichlet, Problem

    m = box_mesh("QUAD9", n, lo=(0, 0), hi=(length, 1.0))
    bcs = [Dirichlet("U", 1, 1.0), Dirichlet("V", 1, 0.0), Dirichlet("V", 2, 0.0),
           Dirichlet("U", 3, 0.0), Dirichlet("V", 3, 0.0), Dirichlet("U", 4, 0.0), Dirichlet("V", 4, 0.0)]
    return Problem(m, rho=1.0, mu=1.0, bcs=bcs)
