"""Pure-Python reader of TF2 tensor-bundle checkpoints (what ``net.load_weights`` reads in
reference model/model.py:262-266), so that ``load_deblender("dc2")`` can use the shipped weights
wherever the checkpoint's data shard exists — no TensorFlow needed.

Format: ``<prefix>.index`` is a LevelDB-style table (blocks of prefix-compressed key/value
entries, 48-byte footer with magic 0xdb4775248b80fb57) whose values are ``BundleEntryProto``
{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}; tensors are raw little-endian in
``<prefix>.data-0000N-of-0000M``.  Blocks of this file family are stored uncompressed.

Every entry carries the masked CRC-32C (Castagnoli) of its bytes; ``load_checkpoint(..., verify=True)`` and
``verify_entry`` check it.  The rule for string tensors (varint lengths, a 4-byte checksum of the lengths, the bytes; the
CRC runs over the lengths as uint32, the length checksum and the bytes) is the one the object-graph entry of the shipped
checkpoint — the only tensor of the one data shard present in the reference snapshot — verifies against
(``tests/test_ckpt_reader.py::test_crc_of_the_real_shard``): offsets, sizes and checksums are read the way TensorFlow
wrote them.
"""
from __future__ import annotations

import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_DT_FLOAT = 1
_DT_STRING = 7


def _crc_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_CRC = _crc_table()


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C (Castagnoli, reflected 0x1EDC6F41) of `data`, continuing from `crc` (crc32c::Extend)."""
    crc ^= 0xFFFFFFFF
    tbl = _CRC
    for x in bytes(data):
        crc = tbl[(crc ^ x) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def crc_mask(c: int) -> int:
    """crc32c::Mask: the form stored in the index (rotate right by 15, add a constant)."""
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


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
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
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
            if f == 6:
                e["crc32c"] = struct.unpack("<I", v[p : p + 4])[0]
            p += 4
        elif wt == 1:
            p += 8
        else:
            raise ValueError(f"unexpected protobuf wire type {wt}")
    return e


def read_index(index_path: str, all_entries: bool = False) -> dict:
    """{key (suffix stripped): entry dict} for the model variables (optimizer slots and the object graph skipped unless
    all_entries)."""
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
            if not k:
                continue  # the bundle header
            if not all_entries and (k == "_CHECKPOINTABLE_OBJECT_GRAPH" or "OPTIMIZER_SLOT" in k or k.startswith("optimizer")):
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


def verify_entry(entry: dict, raw) -> bool:
    """Does `raw` (the entry's bytes in its data shard) carry the checksum the index holds for it?"""
    if entry.get("crc32c") is None:
        return True
    raw = bytes(raw)
    if entry["dtype"] == _DT_STRING:
        # [varint64 length per element][uint32 masked crc of the lengths][bytes]; the entry's crc runs over the lengths as
        # uint32 (uint64 beyond 4 GiB), the length checksum and the bytes (tensor_bundle.cc, WriteStringTensor)
        n = 1
        for d in entry["shape"]:
            n *= d
        p, lens = 0, []
        for _ in range(n):
            ln, p = _varint(raw, p)
            lens.append(ln)
        c = 0
        for ln in lens:
            c = crc32c(struct.pack("<I", ln) if ln < 2**32 else struct.pack("<Q", ln), c)
        stored = struct.unpack("<I", raw[p : p + 4])[0]
        if stored != crc_mask(c):
            return False
        c = crc32c(raw[p : p + 4], c)
        c = crc32c(raw[p + 4 :], c)
        return crc_mask(c) == entry["crc32c"]
    return crc_mask(crc32c(raw)) == entry["crc32c"]


def load_checkpoint(prefix: str, verify: bool = False) -> dict:
    """{key: float32 ndarray} of the model variables of checkpoint `prefix`.  verify=True checks every tensor's CRC-32C
    against the index (pure Python: ~1 MB/s — the DC2 deblender's 33 MB take about half a minute)."""
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
        if verify and not verify_entry(e, raw):
            raise ValueError(f"{path}: CRC-32C mismatch for {key}")
        out[key] = np.frombuffer(bytes(raw), dtype="<f4").reshape(e["shape"]).copy()
    del shards
    return out
