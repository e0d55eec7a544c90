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

    def window_axpy(field_in, stamps, x0, y0, alpha, out=None, field_shape=None, dtype=torch.float64):
        acc = field_in.numpy().copy() if field_in is not None else np.zeros(field_shape)
        view = acc[0] if acc.ndim == 4 else acc
        for s, a, b in zip(stamps.numpy(), x0, y0):
            fo._paste(view, s, int(a), int(b), -1 if alpha < 0 else +1)
        return torch.from_numpy(acc)

    def center_mse(cutouts, means, lo, hi):
        c, m = cutouts.numpy(), means.numpy()
        return torch.from_numpy(np.array([fo.mse(a[lo:hi, lo:hi], b[lo:hi, lo:hi]) for a, b in zip(c, m)]))

    def mse(a, b):
        a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
        b = b.numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
        return float(fo.mse(a, b))

    for name, fn in dict(to_device_field=to_device_field, extract=extract, window_axpy=window_axpy, center_mse=center_mse, mse=mse).items():
        monkeypatch.setattr(_fieldops, name, fn)
