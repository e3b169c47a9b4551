"""CPU restatement of the two O(nnz) passes that follow matrix_fill_full in the Newton loop -- TEST INFRASTRUCTURE.

  row_sum_scale_MSR   src/sl_matrix_util.c:507-600   (called at src/mm_sol_nonlinear.c:1317)
  Loo_norm / L1_norm / L2_norm   src/mm_sol_nonlinear.c:3320-3375, :3275-3296, :3177-3195  (called at :1451-1453
  on the scaled residual, over the owned unknowns)

Pinned: reproduces ``oracle/_ref/goma_ref_fill ... post`` (the reference's own functions) on the committed
fixture tests/golden/post_c1_quad9_ns.npz bit for bit (tests/test_oracle_cpu.py).
"""
import numpy as np


def row_sum_scale_msr(N, a, ija, b):
    """In place on copies; returns (a, b, scale).  ``N`` = owned unknowns (ams->npu)."""
    a, b = a.copy(), b.copy()
    ija = np.asarray(ija, np.int64)
    scale = np.zeros(len(b))
    for irow in range(N):
        k0, k1 = ija[irow], ija[irow + 1]
        row_sum = abs(a[irow])
        for k in range(k0, k1):  # the reference's summation order (sl_matrix_util.c:523-526)
            row_sum += abs(a[k])
        if abs(a[irow]) > 1.0e-200:  # make the diagonal positive (:547-549)
            row_sum = row_sum * (1.0 if a[irow] >= 0 else -1.0)
        scale[irow] = row_sum
        if row_sum == 0.0:  # KEEP_GOING_ON_ZERO_ROW_SUM with WARNING: the divisions below still happen
            pass
        with np.errstate(divide="ignore", invalid="ignore"):
            b[irow] = b[irow] / row_sum
            a[irow] = a[irow] / row_sum
            a[k0:k1] = a[k0:k1] / row_sum
    return a, b, scale


def norms(v, N):
    """(Loo, L1, L2, index of the Loo entry) over the first N entries, single rank."""
    w = np.abs(v[:N])
    k = int(np.argmax(w)) if N else -1
    l2 = 0.0
    for t in v[:N]:  # sequential sum as in L2_norm (:3186-3188)
        l2 += t * t
    l1 = 0.0
    for t in w:
        l1 += t
    return (float(w[k]) if N else -1.0), float(l1), float(np.sqrt(l2)), k


def msr_matvec(N, a, ija, v):
    """w = A v for the owned rows of an MSR matrix: the product of the Newton line search, src/mm_sol_nonlinear.c:442-446
    (``AZ_MSR_matvec_mult``).  Aztec is a third-party dependency that is not vendored in the reference tree (AztecOO of
    Trilinos, ``az_aztec.h``); its published algorithm (Aztec 2.1 user's guide, SAND99-8801J, section 3.1 "DMSR format";
    ``az_matvec_mult.c``) is, per row i, ``val[i] * b[i] + sum_{k = bindx[i]}^{bindx[i+1]-1} val[k] * b[bindx[k]]`` with
    the off-diagonals accumulated in storage order.  Parity unpinned against Aztec itself (no source, no vectors in the
    reference); anchored on the reference's MSR graph and values of the fixtures."""
    ija = np.asarray(ija, np.int64)
    w = np.zeros(len(v))
    for i in range(N):
        acc = 0.0
        for k in range(ija[i], ija[i + 1]):
            acc += a[k] * v[ija[k]]
        w[i] = a[i] * v[i] + acc
    return w

