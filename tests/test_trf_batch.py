"""The batched Trust Region Reflective restatement (debvader_b200/deblend_cutout/trf_batch.py) against
scipy.optimize.least_squares itself — the optimiser the reference calls (deblend_cutout/optimization.py:37-49) — on analytic
objectives shaped like the position fit's: scalar, two variables, bounds +-3, a large constant plus a small multi-modal
part (which local minimum is reached depends on the optimiser's path, so this pins the path), and bowls whose minimum lies
outside the box (the reflective branch)."""
import numpy as np
import pytest
from scipy import optimize

from debvader_b200.deblend_cutout.trf_batch import least_squares_trf_batch


def _family(kind, K, seed):
    rng = np.random.default_rng(seed)
    if kind == "ripples":  # ~0.18 + 1e-5-level ripples: like mean((field - shifted stamp)^2) on a noisy field
        A = rng.uniform(0.5e-5, 3e-5, size=(K, 4))
        W = rng.uniform(0.6, 2.5, size=(K, 4, 2)) * rng.choice([-1, 1], size=(K, 4, 2))
        P = rng.uniform(0, 2 * np.pi, size=(K, 4))
        C = rng.uniform(0.1, 0.3, size=K)
        Q = rng.uniform(0, 2e-6, size=K)

        def f(X, rows):
            # element-wise operations in a fixed order only: a row's value must not depend on what else is in the batch (the
            # 2-point Jacobian divides differences of ~1e-13 by 1.5e-8, so a last-bit change of f moves the path by ~1e-5 px)
            X = np.asarray(X, dtype=np.float64)
            out = C[rows] + Q[rows] * (X[:, 0] * X[:, 0] + X[:, 1] * X[:, 1])
            for j in range(4):
                out = out + A[rows, j] * np.sin(W[rows, j, 0] * X[:, 0] + W[rows, j, 1] * X[:, 1] + P[rows, j])
            return out
    elif kind == "bowls":  # minimum anywhere in [-4.5, 4.5]^2: about half of them outside the box
        M = rng.uniform(-4.5, 4.5, size=(K, 2))
        S = rng.uniform(0.2, 3.0, size=(K, 2))
        R = rng.uniform(-0.6, 0.6, size=K)
        C = rng.uniform(0.05, 2.0, size=K)

        def f(X, rows):
            D = np.asarray(X, dtype=np.float64) - M[rows]
            return C[rows] + 0.01 * (S[rows, 0] * D[:, 0] ** 2 + S[rows, 1] * D[:, 1] ** 2 + R[rows] * np.sqrt(S[rows, 0] * S[rows, 1]) * D[:, 0] * D[:, 1])
    else:  # "zero": residual that reaches 0 (Gauss-Newton regime, gtol / ftol terminations)
        M = rng.uniform(-2.5, 2.5, size=(K, 2))

        def f(X, rows):
            D = np.asarray(X, dtype=np.float64) - M[rows]
            return D[:, 0] * D[:, 0] + D[:, 1] * D[:, 1]
    return f


@pytest.mark.parametrize("kind,K,seed", [("ripples", 60, 1), ("bowls", 60, 2), ("zero", 20, 3), ("ripples", 25, 4)])
def test_batched_trf_walks_scipys_path(kind, K, seed):
    f = _family(kind, K, seed)
    X, info = least_squares_trf_batch(f, K, return_info=True)
    for k in range(K):
        want = optimize.least_squares(lambda x: f(np.asarray(x)[None, :], np.array([k]))[0], (0.0, 0.0), bounds=(-3, 3))
        # same number of evaluations = same sequence of accepted / rejected steps; a bowl whose minimiser sits ON a bound is
        # approached in dozens of ever smaller steps whose count depends on last-bit rounding (LAPACK build of the SVD)
        slack = 0 if kind != "bowls" else max(2, want.nfev // 6)
        assert abs(int(info["nfev"][k]) - want.nfev) <= slack, (kind, k, info["nfev"][k], want.nfev, X[k], want.x)
        # scipy re-evaluates the Jacobian after the terminating step and may then relabel the stop as "gtol" (1): same point
        assert info["status"][k] == want.status or (want.status == 1 and info["status"][k] in (2, 3, 4)), (kind, k, info["status"][k], want.status)
        np.testing.assert_allclose(X[k], want.x, rtol=0, atol=2e-5, err_msg=f"{kind} problem {k}")
    # the point of batching: a handful of evaluation rounds, not one per problem and evaluation
    assert info["rounds"] <= 2 * int(info["nfev"].max()) + 2


def test_batched_trf_empty_and_bounds():
    assert least_squares_trf_batch(lambda X, r: np.zeros(len(r)), 0).shape == (0, 2)
    f = _family("bowls", 40, 9)
    X = least_squares_trf_batch(f, 40)
    assert np.all(np.abs(X) < 3.0)  # strictly feasible, as the reference's optimiser keeps its iterates


def test_batched_trf_on_the_oracle_objective_reproduces_the_reference_fits():
    """The three position fits the REFERENCE itself produced (tests/golden/subpixel.npz: position_optimization run from
    /root/reference by tests/golden/make_golden_subpixel.py), re-run with the batched optimiser around the CPU oracle's
    objective.  Case 1 is the tell-tale: fun has a lower minimum at (0.55, -0.23) than the one scipy's path ends in."""
    import os

    from oracle import spline_numpy as sp

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "subpixel.npz"))
    field, means, dist = g["opt_field"], g["opt_means"], g["opt_dist"]
    F, S = field.shape[1], means.shape[1]
    off = int((F - S) / 2)
    nets = []
    for k in range(len(means)):
        canvas = np.zeros((F, F))
        canvas[off : off + S, off : off + S] = means[k, :, :, 2]
        nets.append(sp.shift_cubic_constant(canvas, dist[k]))

    def fun(X, rows):
        return np.array([sp.position_objective(X[i], field[0, :, :, 2], nets[r]) for i, r in enumerate(rows)])

    x = least_squares_trf_batch(fun, len(means))
    np.testing.assert_allclose(x, g["opt_fitted"], rtol=0, atol=1e-5)
    assert fun(np.array([[0.55, -0.233]]), [1])[0] < fun(x[1:2], [1])[0]  # the multi-modality that makes the path matter
