"""Oracle-backed stand-ins for debvader_b200._fieldops device wrappers, used ONLY by the CPU tests of
the host-side logic (record building, ordering, iteration control).  They never ship."""
import numpy as np
import torch

from oracle import field_numpy as fo


def install(monkeypatch):
    from debvader_b200 import _fieldops

    def to_device_field(field_image, device=None):
        t = field_image if isinstance(field_image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(field_image)))
        return t.contiguous()

    def extract(field_dev, plan, cutout_size, nb_of_bands, out_dtype=torch.float64):
        f = field_dev.numpy()
        n, S = len(plan["ok"]), cutout_size
        out = np.zeros((n, S, S, nb_of_bands))
        idx = []
        for i in range(n):
            if plan["ok"][i] and f.shape[-1] in (nb_of_bands, 1):
                out[i] = f[0, plan["sx"][i] : plan["sx"][i] + plan["lx"][i], plan["sy"][i] : plan["sy"][i] + plan["ly"][i]]
                idx.append(i)
        return torch.from_numpy(out).to(out_dtype), idx

    def window_axpy(field_in, stamps, x0, y0, alpha, out=None, field_shape=None, dtype=torch.float64, planar=False):
        acc = field_in.numpy().copy() if field_in is not None else np.zeros(field_shape)
        view = acc[0] if acc.ndim == 4 else acc
        for s, a, b in zip(stamps.numpy(), x0, y0):
            fo._paste(view, s, int(a), int(b), -1 if alpha < 0 else +1)
        if out is not None:
            out.copy_(torch.from_numpy(acc))
            return out
        return torch.from_numpy(acc)

    def sqdiff_sum_rect(a, b, r0, r1, c0, c1):
        d = a.numpy()[0, r0:r1, c0:c1].astype(np.float64) - b.numpy()[0, r0:r1, c0:c1].astype(np.float64)
        return torch.tensor([float(np.sum(d * d))], dtype=torch.float64)

    def center_mse(cutouts, means, lo, hi):
        c, m = cutouts.numpy(), means.numpy()
        return torch.from_numpy(np.array([fo.mse(a[lo:hi, lo:hi], b[lo:hi, lo:hi]) for a, b in zip(c, m)]))

    def mse(a, b):
        a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
        b = b.numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
        return float(fo.mse(a, b))

    def spline_window_axpy(field_in, stamps, pos_x, pos_y, alpha, field_shape=None, dtype=torch.float64, batch=512, margin=28):
        from oracle import spline_numpy as sp

        if field_in is not None:
            f = field_in.numpy()
            acc = sp.residual_field_subpixel(f, stamps.numpy(), pos_x, pos_y, stamps.shape[1])
            return torch.from_numpy(acc if alpha < 0 else 2 * f - acc)
        return torch.from_numpy(sp.predicted_field_subpixel(field_shape[0], field_shape[2], stamps.numpy(), pos_x, pos_y, stamps.shape[1]))

    for name, fn in dict(to_device_field=to_device_field, extract=extract, window_axpy=window_axpy, center_mse=center_mse, mse=mse,
                         spline_window_axpy=spline_window_axpy, sqdiff_sum_rect=sqdiff_sum_rect).items():
        monkeypatch.setattr(_fieldops, name, fn)


# ---------------------------------------------------------------------------------------------
# numpy emulation of csrc/field_kernels.cu:spline_pass_{x,y}_kernel (the WINDOWED sub-pixel placement),
# used on the CPU to check the kernel's formulation (zero-extension initialisation, geometric tail,
# truncation at +-P) against the full-canvas oracle before any GPU time is spent.
# ---------------------------------------------------------------------------------------------
def _segment(F, S, P, origin):
    lo, hi = max(0, min(F, origin - P)), min(F, max(0, origin + S + P))
    return lo, max(hi, lo)


def _prefilter_lines(lines, lo, hi, F):
    """lines (n, ...): filter along axis 0 like spline_prefilter_line."""
    from oracle.spline_numpy import POLE as z

    c = lines.copy()
    n = hi - lo
    if n < 2:
        return c
    if lo == 0:
        z_n_1 = z ** (F - 1)
        tail = hi == F
        c0 = c[0] + (z_n_1 * c[n - 1] if tail else 0.0)
        z_i = z
        for i in range(1, min(n - 1, F - 2) + 1):
            j = F - 1 - i - lo
            c0 = c0 + z_i * (c[i] + (z_n_1 * c[j] if (tail and 0 <= j < n) else 0.0))
            z_i *= z
        c[0] = c0 / (1.0 - z_n_1 * z_n_1)
    for i in range(1, n):
        c[i] += z * c[i - 1]
    if hi == F:
        c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1.0)
    else:
        c[n - 1] = c[n - 1] * (z / (z * z - 1.0))
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])
    return c


def _eval_lines(c, lo, hi, F, anchor, n_out, pos):
    from oracle.spline_numpy import _mirror_index, spline_weights

    i = anchor + np.arange(n_out)
    cc = i.astype(np.float64) + (-pos)
    valid = ~((cc < 0) | (cc > F - 1))
    start, w = spline_weights(np.where(valid, cc, 0.0))
    out = np.zeros((n_out,) + c.shape[1:])
    bshape = (-1,) + (1,) * (c.ndim - 1)
    for l in range(4):
        idx = _mirror_index(start + l, F)
        ok = (idx >= lo) & (idx < hi)
        v = np.where(ok.reshape(bshape), c[np.clip(idx - lo, 0, max(hi - lo - 1, 0))], 0.0) if hi > lo else 0.0
        out += v * w[:, l].reshape(bshape)
    out[~valid] = 0.0
    return out


def _axis_pass(data, F, P, origin, pos):
    """data (S, ...) along axis 0 at canvas `origin` -> (window (E, ...), anchor)."""
    S = data.shape[0]
    lo, hi = _segment(F, S, P, origin)
    line = np.zeros((hi - lo,) + data.shape[1:])
    for r in range(S):
        u = origin + r - lo
        if 0 <= u < hi - lo:
            line[u] = 6.0 * data[r]
    anchor = origin - P - 1 + int(np.floor(pos))
    return _eval_lines(_prefilter_lines(line, lo, hi, F), lo, hi, F, anchor, S + 2 * P + 2, pos), anchor


def windowed_place(data, px, py, F, P=28, origin=None):
    """(placed (E,E,C) f64, ax, ay) the way dbv_spline_place computes them."""
    S = data.shape[0]
    ox, oy = (int((F - S) / 2),) * 2 if origin is None else origin
    U, ax = _axis_pass(np.asarray(data, dtype=np.float64), F, P, ox, px)  # (E, S, C)
    T, ay = _axis_pass(np.transpose(U, (1, 0, 2)), F, P, oy, py)  # (E_b, E_a, C)
    return np.transpose(T, (1, 0, 2)), ax, ay


def windowed_axpy(field_in, stamps, pos_x, pos_y, alpha, field_shape=None, P=28):
    acc = field_in.copy() if field_in is not None else np.zeros(field_shape)
    view = acc[0] if acc.ndim == 4 else acc
    F = view.shape[0]
    for s, px, py in zip(stamps, pos_x, pos_y):
        T, ax, ay = windowed_place(np.asarray(s), float(px), float(py), F, P)
        E = T.shape[0]
        x0, x1, y0, y1 = max(ax, 0), min(ax + E, F), max(ay, 0), min(ay + E, F)
        if x0 < x1 and y0 < y1:
            view[x0:x1, y0:y1] += alpha * T[x0 - ax : x1 - ax, y0 - ay : y1 - ay]
    return acc


def windowed_objective(r_field, stamp_r, dist, x, P=28):
    """fun of optimization.py:21-33 the way the device path evaluates it: two successive windowed shifts."""
    F = r_field.shape[0]
    T1, a1x, a1y = windowed_place(stamp_r[:, :, None], float(dist[0]), float(dist[1]), F, P)
    T2, ax, ay = windowed_place(T1, float(x[0]), float(x[1]), F, P, origin=(a1x, a1y))
    E = T2.shape[0]
    x0, x1, y0, y1 = max(ax, 0), min(ax + E, F), max(ay, 0), min(ay + E, F)
    t = T2[x0 - ax : x1 - ax, y0 - ay : y1 - ay, 0]
    v = r_field[x0:x1, y0:y1]
    return (np.square(r_field).sum() + (t * t - 2.0 * v * t).sum()) / (F * F)


# ---------------------------------------------------------------------------------------------
# numpy emulation of spw_prefilter (csrc/field_kernels.cu): the cubic prefilter of one line as two warp scans —
# two samples per lane, Kogge-Stone over the 32 lanes with multiplier z^2 — plus the closed-form head and tail.
# ---------------------------------------------------------------------------------------------
def warp_scan_prefilter(x):
    """x: the S <= 64 data samples of a line (gain already applied).  Returns (c[0..S), c[0], c+[S-1]) as the kernel
    computes them (zero state before the data, infinite geometric tail after it)."""
    from oracle.spline_numpy import POLE as z

    S = len(x)
    assert S <= 64
    xs = np.zeros(64)
    xs[:S] = x
    x0, x1 = xs[0::2].copy(), xs[1::2].copy()  # lane l holds samples 2l, 2l+1
    lanes = np.arange(32)
    q = [z**2, z**4, z**8, z**16, z**32]
    kappa = z / (z * z - 1.0)
    # causal
    t0, t1 = x0, x1 + z * x0
    v = t1.copy()
    for d, qd in zip((1, 2, 4, 8, 16), q):
        u = np.concatenate([np.zeros(d), v[:-d]])  # shfl_up
        v = np.where(lanes >= d, v + qd * u, v)
    carry = np.concatenate([[0.0], v[:-1]])
    cp0, cp1 = t0 + z * carry, t1 + z * z * carry
    cplast = (cp1 if (S - 1) & 1 else cp0)[(S - 1) >> 1]
    # anti-causal, c[64] = kappa z c+[63]
    u0, u1 = -z * cp0, -z * cp1
    u1 = u1.copy()
    u1[31] += z * (kappa * z * cp1[31])
    r1 = u1
    r0 = u0 + z * r1
    v = r0.copy()
    for d, qd in zip((1, 2, 4, 8, 16), q):
        u = np.concatenate([v[d:], np.zeros(d)])  # shfl_down
        v = np.where(lanes + d < 32, v + qd * u, v)
    carry = np.concatenate([v[1:], [0.0]])
    c1 = r1 + z * carry
    c0 = r0 + z * z * carry
    c = np.empty(64)
    c[0::2], c[1::2] = c0, c1
    return c[:S], c0[0], cplast
