"""Lazy, device-backed record columns for DeblendField.

The reference returns ``pd.DataFrame(res).to_records(index=False)`` (deblend/field_deblender.py:366-380): a
numpy recarray whose ``cutout_images`` / ``output_images_mean`` / ``output_images_stddev`` /
``epistemic_uncertainty`` columns hold one (S,S,C) ndarray per galaxy.  Materialising those on the host costs
~0.5 MB per galaxy over PCIe (1 GB for a 2000-source field, several times the whole deblending pass), and
``get_residual_field`` only needs them back on the device.  Here the columns hold ``DeviceStamp`` proxies
instead: each behaves like the ndarray it stands for (``np.asarray``, indexing, arithmetic, ``.shape``,
``.dtype``) and fetches its batch from the device — once, all stamps of the column together — the first time a
caller looks at the values.  The record layout (names, order, scalar dtypes) is the reference's.
"""
from __future__ import annotations

import numpy as np
import torch

_NP = {torch.float32: np.float32, torch.float64: np.float64}


class StampBatch:
    """(N,S,S,C) stamps living on the device; ``host()`` downloads them once."""

    def __init__(self, tensor):
        self.tensor = tensor
        self._host = None

    def host(self):
        if self._host is None:
            self._host = self.tensor.detach().cpu().numpy()
        return self._host

    def __len__(self):
        return self.tensor.shape[0]

    def column(self):
        """1-D object array of proxies, one per stamp (what goes into the record column)."""
        n = len(self)
        col = np.empty(n, dtype=object)
        new = DeviceStamp
        for i in range(n):  # element-wise on purpose: numpy must not probe the proxies as sequences (that would download them)
            col[i] = new(self, i)
        return col


class DeviceStamp:
    """One (S,S,C) stamp of a StampBatch; an ndarray as far as numpy is concerned."""

    __slots__ = ("batch", "index")
    __array_priority__ = 100.0

    def __init__(self, batch: StampBatch, index: int):
        self.batch = batch
        self.index = index

    @property
    def tensor(self):
        return self.batch.tensor[self.index]

    def numpy(self):
        return self.batch.host()[self.index]

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        if dtype is not None and a.dtype != dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    @property
    def shape(self):
        return tuple(self.batch.tensor.shape[1:])

    @property
    def dtype(self):
        return np.dtype(_NP[self.batch.tensor.dtype])

    @property
    def ndim(self):
        return self.batch.tensor.dim() - 1

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, k):
        return self.numpy()[k]

    def __iter__(self):
        return iter(self.numpy())

    def __repr__(self):
        return f"DeviceStamp(shape={self.shape}, dtype={self.dtype}, device={self.batch.tensor.device})"

    def astype(self, dtype, **kw):
        return self.numpy().astype(dtype, **kw)

    def copy(self):
        return self.numpy().copy()

    def _bin(name):  # noqa: N805
        def f(self, other):
            return getattr(self.numpy(), name)(np.asarray(other) if isinstance(other, DeviceStamp) else other)

        f.__name__ = name
        return f

    for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__", "__pow__",
               "__lt__", "__le__", "__gt__", "__ge__", "__eq__", "__ne__"):
        locals()[_n] = _bin(_n)
    del _bin, _n
    __hash__ = None

    def __neg__(self):
        return -self.numpy()

    def __getattr__(self, name):  # sum / mean / max / reshape ...: whatever ndarray offers
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.numpy(), name)


def stamp_column(x):
    """record column for a batch of stamps: device tensors become lazy proxies, host arrays stay arrays."""
    if isinstance(x, torch.Tensor):
        return StampBatch(x).column()
    col = np.empty(len(x), dtype=object)
    for i, a in enumerate(x):
        col[i] = a
    return col


def column_batch(records, column):
    """The device tensor behind a record column when the column is exactly the stamps 0..n-1 of ONE StampBatch
    (the records deblend_field just returned), else None."""
    vals = records[column]
    n = len(vals)
    if n == 0 or not isinstance(vals[0], DeviceStamp):
        return None
    b = vals[0].batch
    if len(b) != n:
        return None
    for i in range(n):
        v = vals[i]
        if not isinstance(v, DeviceStamp) or v.batch is not b or v.index != i:
            return None
    return b.tensor


def column_tensor(records, column, device):
    """(n,S,S,C) device tensor of a record column, dtype kept (float32 means / stddevs, float64 epistemic maps and
    user-supplied float64 stamps — field_deblender.py:121-182 pastes them as they are)."""
    t = column_batch(records, column)
    if t is not None:
        return t
    vals = records[column]
    parts = []
    for v in vals:
        if isinstance(v, DeviceStamp):
            parts.append(v.tensor)
        else:
            a = np.asarray(v)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            parts.append(torch.from_numpy(np.ascontiguousarray(a)).to(device))
    dt = torch.float64 if any(p.dtype == torch.float64 for p in parts) else torch.float32
    return torch.stack([p.to(dt) for p in parts]).contiguous()


class quiet_gc:
    """Building the records of a 2000-source field allocates ~20 000 small acyclic objects (three proxies per galaxy, the
    per-galaxy shift arrays ...).  Each 700 of them trigger a generation-0 collection of Python's cyclic garbage collector,
    every hundredth of those a full collection that walks the whole interpreter heap — 37 ms measured on the GPU box with
    torch imported, i.e. eight deblending passes' worth of device time, once every ~10 passes.  None of these objects can
    be part of a cycle, so the collector is paused while they are created and its previous state restored afterwards."""

    def __enter__(self):
        import gc

        self._was = gc.isenabled()
        gc.disable()
        return self

    def __exit__(self, *exc):
        if self._was:
            import gc

            gc.enable()
        return False


def with_quiet_gc(fn):
    """decorator: run `fn` under quiet_gc."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        with quiet_gc():
            return fn(*a, **kw)

    return wrapper


def make_records(columns: dict):
    """``pd.DataFrame(res).to_records(index=False)`` (field_deblender.py:380) built directly with numpy: a recarray
    with the same field names, order and dtypes pandas infers (object for the per-stamp arrays and the shifts, int64 /
    float64 / bool for the scalar columns) — without pandas inspecting, and thereby downloading, the stamp proxies."""
    arrays, names = [], []
    n = None
    for name, v in columns.items():
        if isinstance(v, np.ndarray) and v.dtype == object:
            a = v
        elif isinstance(v, np.ndarray):
            a = v
        elif len(v) and isinstance(v[0], (np.ndarray, list, tuple)):  # e.g. shifts: one small array per galaxy -> object column
            a = np.empty(len(v), dtype=object)
            for i, x in enumerate(v):
                a[i] = x
        else:
            a = np.asarray(v)
            if a.dtype.kind in "US" or a.ndim != 1:
                b = np.empty(len(v), dtype=object)
                for i, x in enumerate(v):
                    b[i] = x
                a = b
        n = len(a) if n is None else n
        if len(a) != n:
            raise ValueError("All arrays must be of the same length")  # pandas' message for ragged columns
        arrays.append(a)
        names.append(name)
    if n == 0:
        dt = [(nm, a.dtype if a.dtype != np.float64 or nm.startswith("galaxy") else object) for nm, a in zip(names, arrays)]
        return np.recarray((0,), dtype=dt)
    return np.rec.fromarrays(arrays, names=names)
