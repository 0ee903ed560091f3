"""Host-side Chebyshev collocation operators for Python callers of kite_colloc_eval (the C++ host has the same in
include/openkite/chebyshev.hpp).  Restates the operator definitions of the reference
(src/kite_math/pseudospectral/chebyshev.hpp:119-232): these are small constant matrices built once on the host and
handed to the GPU evaluator; no dynamics are computed here.
"""
import numpy as np


def colloc_points(P):
    """Chebyshev-Gauss-Lobatto points x_k = cos(k pi / P), k = 0..P (chebyshev.hpp:119-127); index 0 = FINAL time."""
    return np.cos(np.arange(P + 1) * (np.pi / P))


def diff_matrix(P):
    """Trefethen's differentiation matrix with the negative-sum diagonal (chebyshev.hpp:136-153)."""
    n = P + 1
    x = colloc_points(P)
    c = np.where(np.arange(n) % 2 == 1, -1.0, 1.0) * np.where((np.arange(n) == 0) | (np.arange(n) == P), 2.0, 1.0)
    Dn = np.outer(c, 1.0 / c) / ((x[:, None] - x[None, :]) + np.eye(n))
    return Dn - np.diag(Dn.sum(axis=1))


def quad_weights(P):
    """Clenshaw-Curtis weights (chebyshev.hpp:162-195)."""
    theta = np.arange(P + 1) * (np.pi / P)
    w = np.zeros(P + 1)
    v = np.ones(P - 1)
    if P % 2 == 0:
        w[0] = w[P] = 1.0 / (P * P - 1.0)
        for k in range(1, P // 2):
            v -= 2 * np.cos(2 * k * theta[1:P]) / (4.0 * k * k - 1)
        v -= np.cos(P * theta[1:P]) / (P * P - 1.0)
    else:
        w[0] = w[P] = 1.0 / (P * P)
        for k in range(1, (P - 1) // 2 + 1):
            v -= 2 * np.cos(2 * k * theta[1:P]) / (4.0 * k * k - 1)
    w[1:P] = 2 * v / P
    return w


def comp_diff_matrix(P, S):
    """Composite (S segments x order P) block differentiation matrix before the Kronecker product with I
    (chebyshev.hpp:204-229): the last segment carries the full D, earlier segments its first P rows."""
    D = diff_matrix(P)
    if S < 2:
        return D
    n, m = P + 1, S * P + 1
    C = np.zeros((m, m))
    C[m - n:, m - n:] = D
    for k in range(0, (S - 1) * P, P):
        C[k:k + P, k:k + n] = D[:P, :]
    return C
