"""Batched restatement of the optimiser behind the reference's position fit.

Reference call (deblend_cutout/optimization.py:37-49): ``scipy.optimize.least_squares(fun, (0, 0), bounds=(-3, 3))`` with a
SCALAR ``fun`` — i.e. SciPy's Trust Region Reflective algorithm (``method='trf'``, ``tr_solver='exact'``, 2-point Jacobian,
ftol = xtol = gtol = 1e-8, ``max_nfev = 100 * n``) on one residual (m = 1) of two variables (n = 2).  scipy is a third-party
dependency of the reference (``requirements.txt``: scipy==1.11.2; the algorithm is unchanged in the 1.x line); its
published algorithm (Branch, Coleman & Li 1999; Moré 1977 for the trust-region sub-problem) is restated here for MANY
independent problems at once, so that every galaxy of a field walks exactly the path scipy would walk for it while each
round of objective evaluations is ONE batched call (``fun_batch``), not ~40 separate ones per galaxy.

The objective of the position fit is multi-modal at the 1e-5 level (noise), so *which* local minimum is returned depends on
the optimiser's path: a different descent method lands elsewhere (tests/golden/subpixel.npz, case 1).  That is why the
path is restated and not merely "a" bounded minimiser.  ``tests/test_trf_batch.py`` compares it with
``scipy.optimize.least_squares`` itself on analytic multi-modal objectives (CPU).

State machine per problem: OUTER (scaling, termination test, SVD of the augmented Jacobian) -> TRIAL (trust-region step,
one evaluation; repeated with a smaller radius while the cost does not decrease) -> JAC (two evaluations) -> OUTER ...
Each round gathers the pending evaluations of all live problems into one ``fun_batch(X, rows)`` call.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(float).eps
_OUTER, _TRIAL, _JAC, _DONE = 0, 1, 2, 3


def _step_size_to_bound(x, s, lb, ub):
    """smallest t >= 0 with x + t s on a bound, and which coordinates hit (-1 lower / +1 upper / 0)."""
    steps = np.full_like(x, np.inf)
    nz = s != 0
    with np.errstate(over="ignore"):
        steps[nz] = np.maximum((lb - x)[nz] / s[nz], (ub - x)[nz] / s[nz])
    t = steps.min()
    return t, (steps == t) * np.sign(s).astype(int)


def _quad_1d(Jh, g, s, diag, s0=None):
    """coefficients of q(t) = 0.5 (s0 + t s)^T (Jh^T Jh + diag) (s0 + t s) + g^T (s0 + t s); Jh is the single row (m = 1)."""
    v = float(Jh @ s)
    a = 0.5 * (v * v + float(np.dot(s * diag, s)))
    b = float(np.dot(g, s))
    if s0 is None:
        return a, b
    u = float(Jh @ s0)
    b += u * v + float(np.dot(s0 * diag, s))
    c = 0.5 * u * u + float(np.dot(g, s0)) + 0.5 * float(np.dot(s0 * diag, s0))
    return a, b, c


def _min_quad_1d(a, b, lo, hi, c=0.0):
    t = [lo, hi]
    if a != 0:
        e = -0.5 * b / a
        if lo < e < hi:
            t.append(e)
    t = np.asarray(t)
    y = t * (a * t + b) + c
    k = int(np.argmin(y))
    return t[k], y[k]


def _eval_quad(Jh, g, s, diag):
    js = float(Jh @ s)
    return 0.5 * (js * js + float(np.dot(s * diag, s))) + float(np.dot(s, g))


def _select_step_reflective(x, Jh, diag, gh, p, ph, d, Delta, lb, ub, theta):
    """The branch of the step selection taken when the trust-region step leaves the box: the best of (i) the step cut back
    to the interior, (ii) its reflection off the bound it hits first and (iii) the constrained Cauchy step (one problem)."""
    p, ph = p.copy(), ph.copy()
    p_stride, hits = _step_size_to_bound(x, p, lb, ub)
    rh = ph.copy()
    rh[hits.astype(bool)] *= -1
    r = d * rh
    p *= p_stride
    ph *= p_stride
    x_on = x + p
    # intersection of the reflected ray with the trust region
    a = float(np.dot(rh, rh))
    b = float(np.dot(ph, rh))
    c = float(np.dot(ph, ph)) - Delta**2
    disc = np.sqrt(b * b - a * c)
    q = -(b + np.copysign(disc, b))
    t1, t2 = q / a, c / q
    to_tr = max(t1, t2)
    to_bound, _ = _step_size_to_bound(x_on, r, lb, ub)
    r_stride = min(to_bound, to_tr)
    if r_stride > 0:
        r_lo = (1 - theta) * p_stride / r_stride
        r_hi = theta * to_bound if r_stride == to_bound else to_tr
    else:
        r_lo, r_hi = 0, -1
    if r_lo <= r_hi:
        qa, qb, qc = _quad_1d(Jh, gh, rh, diag, s0=ph)
        r_stride, r_value = _min_quad_1d(qa, qb, r_lo, r_hi, c=qc)
        rh = rh * r_stride + ph
        r = rh * d
    else:
        r_value = np.inf
    p *= theta
    ph *= theta
    p_value = _eval_quad(Jh, gh, ph, diag)
    agh = -gh
    ag = d * agh
    to_tr = Delta / np.linalg.norm(agh)
    to_bound, _ = _step_size_to_bound(x, ag, lb, ub)
    ag_stride = theta * to_bound if to_bound < to_tr else to_tr
    qa, qb = _quad_1d(Jh, gh, agh, diag)
    ag_stride, ag_value = _min_quad_1d(qa, qb, 0, ag_stride)
    agh = agh * ag_stride
    ag = ag * ag_stride
    if p_value < r_value and p_value < ag_value:
        return p, ph, -p_value
    if r_value < p_value and r_value < ag_value:
        return r, rh, -r_value
    return ag, agh, -ag_value


def _fd_steps(x, lb, ub):
    """Forward-difference steps of the 2-point Jacobian: sqrt(eps) * sign(x) * max(1, |x|), turned around (or shortened)
    where x + h would leave the box."""
    sign = (x >= 0).astype(float) * 2 - 1
    h = EPS**0.5 * sign * np.maximum(1.0, np.abs(x))
    lower, upper = x - lb, ub - x
    xs = x + h
    violated = (xs < lb) | (xs > ub)
    fitting = np.abs(h) <= np.maximum(lower, upper)
    h = np.where(violated & fitting, -h, h)
    h = np.where((upper >= lower) & ~fitting, upper, h)
    h = np.where((upper < lower) & ~fitting, -lower, h)
    return h


def _strictly_feasible(x, lb, ub):
    xn = x.copy()
    lo = (x - lb) <= np.minimum(ub - x, 0.0)
    hi = (ub - x) <= np.minimum(x - lb, 0.0)
    xn = np.where(lo, np.nextafter(lb, ub), xn)
    xn = np.where(hi, np.nextafter(ub, lb), xn)
    return np.where((xn < lb) | (xn > ub), 0.5 * (lb + ub), xn)


def least_squares_trf_batch(fun_batch, n_problems, x0=(0.0, 0.0), bounds=(-3.0, 3.0), ftol=1e-8, xtol=1e-8, gtol=1e-8,
                            max_nfev=None, return_info=False):
    """``least_squares(fun_k, x0, bounds=bounds)`` (TRF) for k = 0 .. n_problems-1 at once.

    ``fun_batch(X, rows)``: X (m, 2) float64 points, rows (m,) problem indices -> (m,) float64 values of ``fun_rows[i](X[i])``.
    Returns the (n_problems, 2) solutions (and, with return_info, per-problem status / nfev / cost and the number of batched
    evaluation rounds)."""
    K = int(n_problems)
    n = 2
    lb = np.full(n, float(bounds[0]))
    ub = np.full(n, float(bounds[1]))
    x = np.tile(np.asarray(x0, dtype=np.float64).reshape(1, n), (K, 1))
    status = np.zeros(K, dtype=int)
    nfev = np.zeros(K, dtype=int)
    if K == 0:
        return (x, {"status": status, "nfev": nfev, "rounds": 0, "evaluations": 0}) if return_info else x
    if max_nfev is None:
        max_nfev = 100 * n
    allrows = np.arange(K)
    f = np.asarray(fun_batch(x.copy(), allrows), dtype=np.float64).reshape(K)
    nfev += 1
    evaluations, rounds = K, 1
    cost = 0.5 * f * f
    J = np.zeros((K, n))
    g = np.zeros((K, n))
    Delta = np.ones(K)  # norm(x0 / sqrt(v)) = 0 for x0 = 0 -> 1.0; set properly after the first Jacobian
    delta_init = np.ones(K, dtype=bool)
    alpha = np.zeros(K)
    state = np.full(K, _JAC)
    # per-iteration quantities (valid in TRIAL)
    d = np.ones((K, n)); diag_h = np.zeros((K, n)); g_h = np.zeros((K, n)); J_h = np.zeros((K, n))
    uf = np.zeros((K, n)); sv = np.zeros((K, n)); V = np.zeros((K, n, n)); theta = np.ones(K)
    x_new = x.copy(); step_h_norm = np.zeros(K); step_norm = np.zeros(K); pred = np.zeros(K)
    hJ = np.zeros((K, n))

    while True:
        # ---- OUTER: Coleman-Li scaling, first-order optimality, SVD of [J d ; diag(sqrt(g dv))] ------------------------
        idx = np.nonzero(state == _OUTER)[0]
        if idx.size:
            gi, xi = g[idx], x[idx]
            v = np.ones_like(xi); dv = np.zeros_like(xi)
            m = gi < 0
            v[m] = (ub - xi)[m]; dv[m] = -1
            m = gi > 0
            v[m] = (xi - lb)[m]; dv[m] = 1
            first = delta_init[idx]
            if first.any():  # trust radius of the very first iteration
                with np.errstate(divide="ignore", invalid="ignore"):
                    D0 = np.linalg.norm(xi[first] / v[first] ** 0.5, axis=1)
                D0[~(D0 > 0)] = 1.0
                Delta[idx[first]] = D0
                delta_init[idx[first]] = False
            g_norm = np.abs(gi * v).max(axis=1)
            conv = g_norm < gtol
            status[idx[conv]] = 1
            stop = conv | (nfev[idx] >= max_nfev)
            state[idx[stop]] = _DONE
            idx, gi, v, dv, g_norm = idx[~stop], gi[~stop], v[~stop], dv[~stop], g_norm[~stop]
            if idx.size:
                di = v**0.5
                dh = gi * dv
                d[idx], diag_h[idx], g_h[idx] = di, dh, di * gi
                Jh = J[idx] * di
                J_h[idx] = Jh
                A = np.zeros((idx.size, 1 + n, n))
                A[:, 0, :] = Jh
                A[:, 1, 0] = dh[:, 0] ** 0.5
                A[:, 2, 1] = dh[:, 1] ** 0.5
                U, s, Vt = np.linalg.svd(A, full_matrices=False)
                sv[idx] = s
                V[idx] = np.transpose(Vt, (0, 2, 1))
                uf[idx] = U[:, 0, :] * f[idx, None]
                theta[idx] = np.maximum(0.995, 1 - g_norm)
                state[idx] = _TRIAL
        # ---- TRIAL: trust-region sub-problem (m < n: never the Gauss-Newton step), step selection ------------------------
        tri = np.nonzero(state == _TRIAL)[0]
        if tri.size:
            s, Dl = sv[tri], Delta[tri]
            suf = s * uf[tri]
            a_up = np.linalg.norm(suf, axis=1) / Dl
            a_lo = np.zeros(tri.size)
            al = alpha[tri].copy()
            zero = al == 0
            al[zero] = np.maximum(0.001 * a_up[zero], 0.0)
            live = np.ones(tri.size, dtype=bool)
            for _ in range(10):
                if not live.any():
                    break
                reset = live & ((al < a_lo) | (al > a_up))
                al[reset] = np.maximum(0.001 * a_up[reset], (a_lo[reset] * a_up[reset]) ** 0.5)
                with np.errstate(divide="ignore", invalid="ignore"):
                    den = s**2 + al[:, None]
                    p_norm = np.linalg.norm(suf / den, axis=1)
                    phi = p_norm - Dl
                    phi_p = -np.sum(suf**2 / den**3, axis=1) / p_norm
                    ratio = phi / phi_p
                neg = live & (phi < 0)
                a_up[neg] = al[neg]
                a_lo[live] = np.maximum(a_lo[live], (al - ratio)[live])
                al[live] = (al - (phi + Dl) * ratio / Dl)[live]
                live &= ~(np.abs(phi) < 0.01 * Dl)
            with np.errstate(divide="ignore", invalid="ignore"):
                p_h = -np.einsum("kij,kj->ki", V[tri], suf / (s**2 + al[:, None]))
                p_h *= (Dl / np.linalg.norm(p_h, axis=1))[:, None]
            alpha[tri] = al
            p = d[tri] * p_h
            xt = x[tri] + p
            inside = np.all((xt >= lb) & (xt <= ub), axis=1)
            Jp = np.sum(J_h[tri] * p_h, axis=1)
            val = 0.5 * (Jp * Jp + np.sum(p_h * diag_h[tri] * p_h, axis=1)) + np.sum(p_h * g_h[tri], axis=1)
            step, step_h, pr = p.copy(), p_h.copy(), -val
            for j in np.nonzero(~inside)[0]:
                k = tri[j]
                step[j], step_h[j], pr[j] = _select_step_reflective(x[k], J_h[k], diag_h[k], g_h[k], p[j], p_h[j], d[k], Delta[k], lb, ub, theta[k])
            x_new[tri] = _strictly_feasible(x[tri] + step, lb, ub)
            step_h_norm[tri] = np.linalg.norm(step_h, axis=1)
            step_norm[tri] = np.linalg.norm(step, axis=1)
            pred[tri] = pr
        # ---- JAC: forward differences at the accepted point -------------------------------------------------------------
        jac = np.nonzero(state == _JAC)[0]
        if jac.size:
            hJ[jac] = _fd_steps(x[jac], lb, ub)
        if tri.size == 0 and jac.size == 0:
            break
        # ---- one batched evaluation for everything pending ---------------------------------------------------------------
        xj0 = x[jac].copy(); xj0[:, 0] = x[jac, 0] + hJ[jac, 0]
        xj1 = x[jac].copy(); xj1[:, 1] = x[jac, 1] + hJ[jac, 1]
        X = np.concatenate([x_new[tri], xj0, xj1])
        rows = np.concatenate([tri, jac, jac])
        vals = np.asarray(fun_batch(X, rows), dtype=np.float64).reshape(-1)
        evaluations += rows.size
        rounds += 1
        ft, fj0, fj1 = vals[: tri.size], vals[tri.size : tri.size + jac.size], vals[tri.size + jac.size :]
        # ---- TRIAL answers ---------------------------------------------------------------------------------------------------
        if tri.size:
            nfev[tri] += 1
            finite = np.isfinite(ft)
            bad = tri[~finite]
            Delta[bad] = 0.25 * step_h_norm[bad]  # retried (or given up when nfev is exhausted), alpha kept
            gave_up = bad[nfev[bad] >= max_nfev]
            state[gave_up] = _DONE
            ok, fo = tri[finite], ft[finite]
            if ok.size:
                cost_new = 0.5 * fo * fo
                actual = cost[ok] - cost_new
                pr, shn, Dl = pred[ok], step_h_norm[ok], Delta[ok]
                with np.errstate(divide="ignore", invalid="ignore"):
                    ratio = np.where(pr > 0, actual / pr, np.where((pr == 0) & (actual == 0), 1.0, 0.0))
                D_new = np.where(ratio < 0.25, 0.25 * shn, np.where((ratio > 0.75) & (shn > 0.95 * Dl), 2.0 * Dl, Dl))
                ft_ok = (actual < ftol * cost[ok]) & (ratio > 0.25)
                xt_ok = step_norm[ok] < xtol * (xtol + np.linalg.norm(x[ok], axis=1))
                term = np.where(ft_ok & xt_ok, 4, np.where(ft_ok, 2, np.where(xt_ok, 3, 0)))
                cont = term == 0
                with np.errstate(divide="ignore", invalid="ignore"):
                    alpha[ok[cont]] = (alpha[ok] * Dl / D_new)[cont]
                Delta[ok[cont]] = D_new[cont]
                status[ok[~cont]] = term[~cont]
                acc = actual > 0
                ka = ok[acc]
                x[ka] = x_new[ka]
                f[ka] = fo[acc]
                cost[ka] = cost_new[acc]
                # accepted and still running -> Jacobian; terminated -> done (x already holds the accepted point);
                # rejected -> another trial with the smaller radius unless the evaluation budget is spent
                state[ok[acc & cont]] = _JAC
                state[ok[~cont]] = _DONE
                rej = ok[~acc & cont]
                state[rej[nfev[rej] >= max_nfev]] = _DONE
        # ---- JAC answers ---------------------------------------------------------------------------------------------------
        if jac.size:
            dx0 = (x[jac, 0] + hJ[jac, 0]) - x[jac, 0]
            dx1 = (x[jac, 1] + hJ[jac, 1]) - x[jac, 1]
            J[jac, 0] = (fj0 - f[jac]) / dx0
            J[jac, 1] = (fj1 - f[jac]) / dx1
            g[jac] = J[jac] * f[jac, None]
            state[jac] = _OUTER
    if return_info:
        return x, {"status": status, "nfev": nfev, "cost": cost, "rounds": rounds, "evaluations": evaluations}
    return x
