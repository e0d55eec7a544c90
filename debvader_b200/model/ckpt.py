"""Pure-Python reader of TF2 tensor-bundle checkpoints (what ``net.load_weights`` reads in
reference model/model.py:262-266), so that ``load_deblender("dc2")`` can use the shipped weights
wherever the checkpoint's data shard exists — no TensorFlow needed.

Format: ``<prefix>.index`` is a LevelDB-style table (blocks of prefix-compressed key/value
entries, 48-byte footer with magic 0xdb4775248b80fb57) whose values are ``BundleEntryProto``
{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}; tensors are raw little-endian in
``<prefix>.data-0000N-of-0000M``.  Blocks of this file family are stored uncompressed.
"""
from __future__ import annotations

import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_DT_FLOAT = 1


def _varint(buf, p):
    r = s = 0
    while True:
        b = buf[p]
        p += 1
        r |= (b & 0x7F) << s
        if not b & 0x80:
            return r, p
        s += 7


def _block_entries(data, off, size):
    blk = data[off : off + size]
    if data[off + size] != 0:
        raise ValueError("compressed checkpoint index blocks are not supported")
    nrestart = struct.unpack("<I", blk[-4:])[0]
    end = len(blk) - 4 - 4 * nrestart
    p, key = 0, b""
    while p < end:
        shared, p = _varint(blk, p)
        non, p = _varint(blk, p)
        vlen, p = _varint(blk, p)
        key = key[:shared] + blk[p : p + non]
        p += non
        yield key, blk[p : p + vlen]
        p += vlen


def _parse_shape(sub):
    shape, q = [], 0
    while q < len(sub):
        tag, q = _varint(sub, q)
        if tag & 7 == 2:
            ln, q = _varint(sub, q)
            dim = sub[q : q + ln]
            q += ln
            if tag >> 3 == 2:  # TensorShapeProto.dim
                r = 0
                size = 0
                while r < len(dim):
                    t3, r = _varint(dim, r)
                    if t3 & 7 == 0:
                        val, r = _varint(dim, r)
                        if t3 >> 3 == 1:
                            size = val
                    else:
                        l3, r = _varint(dim, r)
                        r += l3
                shape.append(size)
        else:
            _, q = _varint(sub, q)
    return shape


def _parse_entry(v):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0}
    p = 0
    while p < len(v):
        tag, p = _varint(v, p)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            val, p = _varint(v, p)
            if f == 1:
                e["dtype"] = val
            elif f == 3:
                e["shard_id"] = val
            elif f == 4:
                e["offset"] = val
            elif f == 5:
                e["size"] = val
        elif wt == 2:
            ln, p = _varint(v, p)
            if f == 2:
                e["shape"] = _parse_shape(v[p : p + ln])
            p += ln
        elif wt == 5:
            p += 4
        elif wt == 1:
            p += 8
        else:
            raise ValueError(f"unexpected protobuf wire type {wt}")
    return e


def read_index(index_path: str) -> dict:
    """{key (suffix stripped): entry dict} for the model variables (optimizer slots skipped)."""
    data = open(index_path, "rb").read()
    footer = data[-48:]
    if struct.unpack("<Q", footer[-8:])[0] != _MAGIC:
        raise ValueError(f"{index_path}: not a tensor-bundle index (bad magic)")
    p = 0
    _, p = _varint(footer, p)
    _, p = _varint(footer, p)
    ioff, p = _varint(footer, p)
    isz, p = _varint(footer, p)
    out = {}
    for _, handle in _block_entries(data, ioff, isz):
        boff, q = _varint(handle, 0)
        bsz, q = _varint(handle, q)
        for key, val in _block_entries(data, boff, bsz):
            k = key.decode()
            if not k or k == "_CHECKPOINTABLE_OBJECT_GRAPH" or "OPTIMIZER_SLOT" in k or k.startswith("optimizer"):
                continue
            if k.endswith(_SUFFIX):
                k = k[: -len(_SUFFIX)]
            out[k] = _parse_entry(val)
    return out


def latest_checkpoint(directory: str):
    """tf.train.latest_checkpoint: reads the text file `checkpoint` (model/model.py:265)."""
    path = os.path.join(directory, "checkpoint")
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("model_checkpoint_path:"):
            name = line.split(":", 1)[1].strip().strip('"')
            return name if os.path.isabs(name) else os.path.join(directory, name)
    return None


def load_checkpoint(prefix: str) -> dict:
    """{key: float32 ndarray} of the model variables of checkpoint `prefix`."""
    entries = read_index(prefix + ".index")
    shards = sorted({e["shard_id"] for e in entries.values()})
    import glob

    files = sorted(glob.glob(prefix + ".data-*-of-*"))
    if not files:
        raise FileNotFoundError(f"no data shard found for {prefix}")
    n_of = int(files[0].rsplit("-of-", 1)[1])
    out = {}
    handles = {}
    for key, e in entries.items():
        if e["dtype"] != _DT_FLOAT:
            continue
        path = f"{prefix}.data-{e['shard_id']:05d}-of-{n_of:05d}"
        if path not in handles:
            if not os.path.exists(path) or os.path.getsize(path) < e["offset"] + e["size"]:
                raise FileNotFoundError(
                    f"checkpoint data shard {path} is missing or truncated; the tensor data of the DC2 "
                    "deblender is not part of this snapshot (see load_deblender(weights=...))"
                )
            handles[path] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = handles[path][e["offset"] : e["offset"] + e["size"]]
        out[key] = np.frombuffer(bytes(raw), dtype="<f4").reshape(e["shape"]).copy()
    del shards
    return out
