"""CPU oracle of the detection step — TEST INFRASTRUCTURE, never imported by the product.

reference: detect/detection.py:5-56 calls the third-party C library `sep` (unpinned in requirements.txt:8; absent from
/root/reference and not installable here):

    bkg = sep.Background(r_band)                      detection.py:15   (defaults: 64x64 meshes, 3x3 median filter)
    sep.extract(r_band - bkg, thresh=1.5, err=bkg.globalrms, minarea=4, filter_kernel=<7x7>, filter_type="conv",
                deblend_nthresh=64, deblend_cont=1e-5)                   detection.py:37-46
    centres = (round(y - int(F/2)), round(x - int(F/2)))                detection.py:48-54

**PARITY UNPINNED**: `sep` cannot be run here, the reference holds no golden detections, so this file restates the
PUBLISHED algorithm (Bertin & Arnouts 1996, SExtractor; sep = Barbary 2016) stage by stage in its own words:

  B1  per 64x64 mesh: mean / sigma, one clip at +-2 sigma, a histogram of <= 4096 levels over +-5 sigma, iterated 3-sigma
      clipping around the histogram median; background = 2.5 median - 1.5 mean when |mean - median| < 0.3 sigma, else the median
  B2  bad meshes filled from the nearest good ones, 3x3 median filter of the mesh maps, global background / rms = medians
  B3  bicubic-spline interpolation of the mesh map (natural splines over mesh centres, y then x, extrapolated at the rims)
  E1  matched filter: the 7x7 mask normalised by the sum of its absolute values, zero outside the image
  E2  threshold 1.5 x global rms on the FILTERED image, 8-connected components of >= 4 pixels
  E3  object order = the order in which Lutz's one-pass scan COMPLETES objects = ascending (last row, last column on that row)
      = ascending largest raster index of the object's pixels
  E4  barycentre of the UNFILTERED foreground over the object's pixels (double precision), centres rounded half-to-even

NOT restated (documented gaps): the multi-threshold deblending of composite objects (64 levels, contrast 1e-5)
and the `clean` pass of sep.extract — a connected footprint stays ONE detection here; the iterative loop
(deblend_iterative/iterative_deblender.py:21-99) finds the remaining members of a blend in the residual of the next step.

Every arithmetic step is written in an explicit order and precision (float32 where sep's PIXTYPE is float, float64 sums in a
fixed sequence) so that the CUDA detector (csrc/detect_kernels.cu) can be compared with this file BIT FOR BIT.
"""
import numpy as np

BW = 64  # mesh size (sep.Background defaults bw = bh = 64)
NSIGMA, AMIN, MAXLEVELS = 5.0, 4.0, 4096  # histogram quantisation
MINGOODFRAC = 0.5
BIG = np.float32(1e30)
DETECT_THRESH = 1.5
MINAREA = 4

f32 = np.float32
f64 = np.float64


def _seq_sum_mesh(a, valid):
    """(ny, BW, nx, BW) float64 -> per-mesh sums in the order: the rows of a mesh column one after the other, then the columns."""
    a = np.where(valid, a, 0.0)
    col = np.zeros((a.shape[0], a.shape[2], a.shape[3]), f64)  # (ny, nx, BW)
    for r in range(BW):
        col = col + a[:, r, :, :]
    tot = np.zeros((a.shape[0], a.shape[2]), f64)
    for c in range(BW):
        tot = tot + col[:, :, c]
    return tot


def _meshes(img):
    h, w = img.shape
    ny, nx = (h - 1) // BW + 1, (w - 1) // BW + 1
    pad = np.zeros((ny * BW, nx * BW), img.dtype)
    pad[:h, :w] = img
    valid = np.zeros((ny * BW, nx * BW), bool)
    valid[:h, :w] = True
    return pad.reshape(ny, BW, nx, BW), valid.reshape(ny, BW, nx, BW), ny, nx


def histogram_guess(histo, nlevels, mean0, sigma0, qzero, qscale):
    """B1, second half: iterated clipping on the level histogram of ONE mesh (doubles; counts are integers, so the sums are
    exact).  Returns (background, sigma) as float32."""
    nm1 = nlevels - 1
    lcut, hcut = 0, nm1
    sig, sig1 = 10.0 * nm1, 1.0
    mea = med = f64(mean0)
    n = 100
    while n > 0 and sig >= 0.1 and abs(sig / sig1 - 1.0) > 1e-4:
        n -= 1
        sig1 = sig
        tot = 0
        mea = 0.0
        sig = 0.0
        lowsum = highsum = 0
        lo, hi = lcut, hcut
        for i in range(lcut, hcut + 1):
            if lowsum < highsum:
                lowsum += int(histo[lo])
                lo += 1
            else:
                highsum += int(histo[hi])
                hi -= 1
            c = int(histo[i])
            tot += c
            mea += f64(c) * f64(i)
            sig += f64(c) * f64(i) * f64(i)
        if hi >= 0:
            big = max(int(histo[min(lo, nm1)]), int(histo[hi]))
            med = f64(hi) + 0.5 + (f64(highsum - lowsum) / (2.0 * f64(big)) if big > 0 else 0.0)
        else:
            med = f64(0.0)
        if tot:
            mea = mea / f64(tot)
            sig = sig / f64(tot) - mea * mea
        sig = np.sqrt(sig) if sig > 0.0 else f64(0.0)
        t = med - 3.0 * sig
        lcut = int(t + 0.5) if t > 0.0 else 0
        t = med + 3.0 * sig
        hcut = (int(t + 0.5) if t > 0.0 else int(t - 0.5)) if t < nm1 else nm1
        if hcut < lcut:
            hcut = lcut
    if sig > 0.0:
        if abs((mea - med) / sig) < 0.3:
            back = f64(qzero) + (2.5 * med - 1.5 * mea) * f64(qscale)
        else:
            back = f64(qzero) + med * f64(qscale)
    else:
        back = f64(qzero) + mea * f64(qscale)
    return f32(back), f32(sig * f64(qscale))


def mesh_statistics(img32):
    """B1 for every mesh: (back, sigma) float32 maps of shape (ny, nx); bad meshes hold -BIG."""
    m, valid, ny, nx = _meshes(img32)
    m64 = m.astype(f64)
    npix_all = _seq_sum_mesh(valid.astype(f64), valid)
    s1 = _seq_sum_mesh(m64, valid)
    s2 = _seq_sum_mesh(m64 * m64, valid)
    mean = s1 / npix_all
    var = s2 / npix_all - mean * mean
    sigma = np.where(var > 0.0, np.sqrt(np.maximum(var, 0.0)), 0.0)
    lcut = (mean - 2.0 * sigma).astype(f32)
    hcut = (mean + 2.0 * sigma).astype(f32)
    keep = valid & (m >= lcut[:, None, :, None]) & (m <= hcut[:, None, :, None])
    npix = _seq_sum_mesh(keep.astype(f64), keep)
    t1 = _seq_sum_mesh(m64, keep)
    t2 = _seq_sum_mesh(m64 * m64, keep)
    back = np.full((ny, nx), -BIG, f32)
    sig = np.full((ny, nx), -BIG, f32)
    step = np.sqrt(2.0 / np.pi) * NSIGMA / AMIN
    for j in range(ny):
        for i in range(nx):
            n = npix[j, i]
            if n < npix_all[j, i] * MINGOODFRAC or n < 1:
                continue
            mean2 = t1[j, i] / n
            var2 = t2[j, i] / n - mean2 * mean2
            sigma2 = np.sqrt(var2) if var2 > 0.0 else f64(0.0)
            nlevels = min(int(step * n + 1.0), MAXLEVELS)
            qscale = f32(2.0 * NSIGMA * sigma2 / f64(nlevels)) if sigma2 > 0.0 else f32(1.0)
            qzero = f32(mean2 - NSIGMA * sigma2)
            cste = f32(0.499999 - f64(qzero) / f64(qscale))
            v = m[j, :, i, :][valid[j, :, i, :]]
            # level of a pixel: (int)(pix / qscale + cste) in float32, truncation toward zero
            lev = ((v / qscale).astype(f32) + cste).astype(f32)
            ok = (lev > f32(-1.0)) & (lev < f32(nlevels))
            b = np.trunc(lev[ok]).astype(np.int64)
            b = b[(b >= 0) & (b < nlevels)]
            histo = np.bincount(b, minlength=nlevels)
            back[j, i], sig[j, i] = histogram_guess(histo, nlevels, mean2, sigma2, qzero, qscale)
    return back, sig


def _median_f32(values):
    """median as sep's fqmedian: middle element, or the float32 mean of the two middle ones"""
    a = np.sort(np.asarray(values, f32))
    n = len(a)
    if n % 2:
        return a[n // 2]
    return f32(f32(a[n // 2 - 1] + a[n // 2]) * f32(0.5))


def filter_meshes(back, sig):
    """B2: bad-mesh fill, 3x3 median filter, global medians.  Returns (back, sigma, globalback, globalrms)."""
    ny, nx = back.shape
    back, sig = back.copy(), sig.copy()
    bad = back <= -BIG
    if bad.any() and not bad.all():
        gy, gx = np.nonzero(~bad)
        b0, s0 = back.copy(), sig.copy()
        for y, x in zip(*np.nonzero(bad)):
            d2 = (gy - y) ** 2 + (gx - x) ** 2
            near = d2 == d2.min()
            # float32 running sums in raster order of the good meshes
            sb, ss, k = f32(0), f32(0), 0
            for yy, xx in zip(gy[near], gx[near]):
                sb = f32(sb + b0[yy, xx])
                ss = f32(ss + s0[yy, xx])
                k += 1
            back[y, x] = f32(sb / f32(k))
            sig[y, x] = f32(ss / f32(k))
    fb, fs = back.copy(), sig.copy()
    for y in range(ny):
        for x in range(nx):
            ys, xs = slice(max(y - 1, 0), min(y + 2, ny)), slice(max(x - 1, 0), min(x + 2, nx))
            fb[y, x] = _median_f32(back[ys, xs].ravel())
            fs[y, x] = _median_f32(sig[ys, xs].ravel())
    return fb, fs, _median_f32(fb.ravel()), _median_f32(fs.ravel())


def _spline_d2(v):
    """second derivatives / 6 of the natural cubic spline through v (unit spacing) along axis 0, float32 Thomas algorithm:
    M[k-1] + 4 M[k] + M[k+1] = 6 (v[k+1] - 2 v[k] + v[k-1]), M[0] = M[n-1] = 0; returns M / 6."""
    v = np.asarray(v, f32)
    n = v.shape[0]
    d = np.zeros_like(v)
    if n < 3:
        return d
    cp = np.zeros_like(v)  # modified super-diagonal
    u = np.zeros_like(v)  # modified right-hand side
    for k in range(1, n - 1):
        rhs = f32(6.0) * (f32(v[k + 1] + v[k - 1]) - f32(f32(2.0) * v[k]))
        den = f32(4.0) - cp[k - 1]
        cp[k] = f32(1.0) / den
        u[k] = (rhs - u[k - 1]) / den
    m = np.zeros_like(v)
    for k in range(n - 2, 0, -1):
        m[k] = u[k] - cp[k] * m[k + 1]
    return (m / f32(6.0)).astype(f32)


def _spline_eval(lo, hi, dlo, dhi, t):
    ct = f32(1.0) - t
    return f32(ct * lo) + f32(t * hi) + f32(f32(f32(f32(ct * ct) * ct) - ct) * dlo) + f32(f32(f32(f32(t * t) * t) - t) * dhi)


def background_map(back, h, w):
    """B3: float32 (h, w) map from the (ny, nx) mesh map."""
    ny, nx = back.shape
    dback = _spline_d2(back)  # along y, per mesh column
    ys = np.arange(h, dtype=f32)
    if ny > 1:
        u = ((ys + f32(0.5)) / f32(BW)).astype(f32) - f32(0.5)
        yl = np.clip(np.floor(u).astype(np.int64), 0, ny - 2)
        t = (u - yl.astype(f32)).astype(f32)[:, None]
        node = _spline_eval(back[yl], back[yl + 1], dback[yl], dback[yl + 1], t).astype(f32)  # (h, nx)
    else:
        node = np.repeat(back[:1], h, axis=0)
    dnode = _spline_d2(node.T).T if nx > 1 else np.zeros_like(node)
    xs = np.arange(w, dtype=f32)
    if nx > 1:
        u = ((xs + f32(0.5)) / f32(BW)).astype(f32) - f32(0.5)
        xl = np.clip(np.floor(u).astype(np.int64), 0, nx - 2)
        t = (u - xl.astype(f32)).astype(f32)[None, :]
        return _spline_eval(node[:, xl], node[:, xl + 1], dnode[:, xl], dnode[:, xl + 1], t).astype(f32)
    return np.repeat(node[:, :1], w, axis=1)


def normalised_filter(kernel):
    """E1: float32 taps divided by the float32 running sum of their absolute values (raster order)."""
    k = np.asarray(kernel, f32)
    s = f32(0.0)
    for v in k.ravel():
        s = f32(s + abs(v))
    return (k / s).astype(f32)


def matched_filter(img32, taps):
    """E1: out[y, x] = sum over (ky, kx) in raster order of taps[ky, kx] * img[y + ky - 3, x + kx - 3], float32, taps outside
    the image skipped."""
    h, w = img32.shape
    kh, kw = taps.shape
    out = np.zeros((h, w), f32)
    for ky in range(kh):
        for kx in range(kw):
            dy, dx = ky - kh // 2, kx - kw // 2
            y0, y1 = max(0, -dy), min(h, h - dy)
            x0, x1 = max(0, -dx), min(w, w - dx)
            if y0 >= y1 or x0 >= x1:
                continue
            out[y0:y1, x0:x1] = out[y0:y1, x0:x1] + (taps[ky, kx] * img32[y0 + dy : y1 + dy, x0 + dx : x1 + dx]).astype(f32)
    return out


def extract_objects(fg, taps, thresh, minarea=MINAREA):
    """E1-E4 on a float32 foreground image: returns (conv, [(last raster index, x, y, npix), ...] in detection order)."""
    from scipy import ndimage

    h, w = fg.shape
    conv = matched_filter(fg, taps)
    mask = conv > thresh
    lab, n = ndimage.label(mask, structure=np.ones((3, 3), int))
    objs = []
    if n:
        idx = np.flatnonzero(mask)
        li = lab.ravel()[idx]
        order = np.argsort(li, kind="stable")
        idx, li = idx[order], li[order]
        starts = np.flatnonzero(np.r_[True, li[1:] != li[:-1]])
        ends = np.r_[starts[1:], len(li)]
        fgr, cvr = fg.ravel(), conv.ravel()
        for s, e in zip(starts, ends):
            p = idx[s:e]  # raster order
            if len(p) < minarea:
                continue
            yy, xx = p // w, p % w
            xmin, ymin = xx.min(), yy.min()
            val = fgr[p].astype(f64)
            tv = np.cumsum(val)[-1]
            if not tv > 0.0:  # the unfiltered flux of a faint detection can be <= 0: weight with the filtered values (> thresh > 0)
                val = cvr[p].astype(f64)
                tv = np.cumsum(val)[-1]
            mx = np.cumsum(val * (xx - xmin).astype(f64))[-1]
            my = np.cumsum(val * (yy - ymin).astype(f64))[-1]
            objs.append((int(p.max()), mx / tv + xmin, my / tv + ymin, len(p)))
    objs.sort(key=lambda o: o[0])
    return conv, objs


def centres_of(objs, H, W):
    """detection.py:48-54: (row, col) offsets from int(F/2), np.round (half to even)"""
    x = np.array([o[1] for o in objs], f64)
    y = np.array([o[2] for o in objs], f64)
    c = np.stack([np.round(y - int(H / 2)), np.round(x - int(W / 2))], axis=1) if len(objs) else np.zeros((0, 2), f64)
    return c, x, y


def detect(field_image, kernel, band=2, return_details=False):
    """The whole step: returns the (N, 2) float64 array detect_objects returns — (row, col) offsets from the field centre —
    in detection order."""
    field_image = np.asarray(field_image)
    r64 = field_image[0, :, :, band].astype(f64)
    r32 = r64.astype(f32)
    h, w = r32.shape
    back0, sig0 = mesh_statistics(r32)
    back, sig, gback, grms = filter_meshes(back0, sig0)
    bmap = background_map(back, h, w)
    fg = (r64 - bmap.astype(f64)).astype(f32)
    taps = normalised_filter(kernel)
    thresh = f32(f64(DETECT_THRESH) * f64(grms))
    conv, objs = extract_objects(fg, taps, thresh)
    centres, x, y = centres_of(objs, h, w)
    if return_details:
        return centres, {"x": x, "y": y, "npix": np.array([o[3] for o in objs], np.int64), "last": np.array([o[0] for o in objs], np.int64),
                         "back": back, "sigma": sig, "back_raw": back0, "sigma_raw": sig0, "globalback": gback, "globalrms": grms,
                         "bmap": bmap, "fg": fg, "conv": conv, "thresh": thresh, "taps": taps}
    return centres
