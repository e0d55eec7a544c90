"""Pure-Python TF tensor-bundle reader: round trip on a synthetic bundle written in the same format,
and the architecture golden parsed from the shipped DC2 index (tests/golden/architecture.json)."""
import json
import os
import struct

import numpy as np
import pytest

from debvader_b200.model import ckpt, spec


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _block(entries):
    # no prefix compression, a single restart point
    body = b"".join(_varint(0) + _varint(len(k)) + _varint(len(v)) + k + v for k, v in entries)
    return body + struct.pack("<I", 0) + struct.pack("<I", 1)


def _entry(shape, offset, size, shard=0):
    dims = b"".join(b"\x12" + _varint(len(d)) + d for d in (b"\x08" + _varint(s) for s in shape))
    return b"\x08\x01" + b"\x12" + _varint(len(dims)) + dims + b"\x18" + _varint(shard) + b"\x20" + _varint(offset) + b"\x28" + _varint(size)


def write_bundle(prefix, tensors):
    data = bytearray()
    entries = []
    for key in sorted(tensors):
        a = np.ascontiguousarray(tensors[key], dtype="<f4")
        entries.append(((key + "/.ATTRIBUTES/VARIABLE_VALUE").encode(), _entry(a.shape, len(data), a.nbytes)))
        data += a.tobytes()
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    blk = _block(entries)
    trailer = b"\x00" + b"\x00\x00\x00\x00"
    f = bytearray(blk + trailer)
    handle = _varint(0) + _varint(len(blk))
    idx = _block([(entries[-1][0], handle)])
    idx_off = len(f)
    f += idx + trailer
    meta = _block([])
    meta_off = len(f)
    f += meta + trailer
    footer = _varint(meta_off) + _varint(len(meta)) + _varint(idx_off) + _varint(len(idx))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    f += footer
    open(prefix + ".index", "wb").write(bytes(f))
    open(os.path.join(os.path.dirname(prefix), "checkpoint"), "w").write(f'model_checkpoint_path: "{os.path.basename(prefix)}"\n')


def test_bundle_round_trip(tmp_path):
    w = spec.random_weights(seed=3)
    small = {k: v for k, v in w.items() if v.size < 40000}
    prefix = str(tmp_path / "w.ckpt")
    write_bundle(prefix, small)
    assert ckpt.latest_checkpoint(str(tmp_path)) == prefix
    got = ckpt.load_checkpoint(prefix)
    assert set(got) == set(small)
    for k in small:
        np.testing.assert_array_equal(got[k], small[k])


def test_missing_data_shard_is_a_clear_error(tmp_path):
    prefix = str(tmp_path / "w.ckpt")
    write_bundle(prefix, {"layer_with_weights-0/layer_with_weights-0/gamma": np.ones(6, np.float32)})
    os.remove(prefix + ".data-00000-of-00001")
    with pytest.raises(FileNotFoundError):
        ckpt.load_checkpoint(prefix)


def test_spec_matches_shipped_checkpoint_index(golden_dir):
    a = json.load(open(os.path.join(golden_dir, "architecture.json")))
    table = spec.tensor_table()
    assert len(table) == 64
    for key, shape in table:
        assert a["tensors"][key] == list(shape), key
    w = spec.random_weights(1)
    enc = sum(v.size for k, v in w.items() if k.startswith("layer_with_weights-0/"))
    dec = sum(v.size for k, v in w.items() if k.startswith("layer_with_weights-1/"))
    assert {"encoder": enc, "decoder": dec, "total": enc + dec} == a["net_summary_params"]


@pytest.mark.skipif(not os.path.exists("/root/reference/src/debvader/data/weights/dc2/checkpoint"), reason="build container only")
def test_reader_on_the_real_index():
    d = "/root/reference/src/debvader/data/weights/dc2"
    latest = ckpt.latest_checkpoint(d)
    entries = ckpt.read_index(latest + ".index")
    assert {k: tuple(e["shape"]) for k, e in entries.items()} == dict(spec.tensor_table())
    with pytest.raises(FileNotFoundError):  # the 99.8 MB data shard is not part of the snapshot
        ckpt.load_checkpoint(latest)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/debvader/data/weights/dc2/checkpoint"), reason="build container only")
def test_crc_of_the_real_shard():
    """The one data shard the reference snapshot does contain (…data-00000-of-00002: the object graph, a string tensor): its
    bytes, located through the index the way the reader locates every tensor, carry exactly the CRC-32C TensorFlow stored
    for them — offsets, sizes, the string-tensor framing and the checksum are read as they were written."""
    d = "/root/reference/src/debvader/data/weights/dc2"
    prefix = ckpt.latest_checkpoint(d)
    entries = ckpt.read_index(prefix + ".index", all_entries=True)
    in_shard0 = {k: e for k, e in entries.items() if e["shard_id"] == 0 and e["size"] > 0}
    assert list(in_shard0) == ["_CHECKPOINTABLE_OBJECT_GRAPH"]
    e = in_shard0["_CHECKPOINTABLE_OBJECT_GRAPH"]
    raw = open(prefix + ".data-00000-of-00002", "rb").read()
    assert e["offset"] + e["size"] == len(raw) and e["crc32c"] is not None
    assert ckpt.verify_entry(e, raw[e["offset"] : e["offset"] + e["size"]])
    bad = bytearray(raw)
    bad[100] ^= 1
    assert not ckpt.verify_entry(e, bytes(bad)[e["offset"] : e["offset"] + e["size"]])
    assert all(v["crc32c"] is not None for v in ckpt.read_index(prefix + ".index").values())  # every model tensor has one too


def test_crc32c_known_answers_and_float_tensor_round_trip(tmp_path):
    """CRC-32C check values (RFC 3720 appendix B.4) and verify=True on a checkpoint written by this module's test writer."""
    assert ckpt.crc32c(b"123456789") == 0xE3069283
    assert ckpt.crc32c(bytes(32)) == 0x8A9136AA
    assert ckpt.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert ckpt.crc32c(b"6789", ckpt.crc32c(b"12345")) == 0xE3069283  # Extend
    a = np.arange(12, dtype="<f4").reshape(3, 4)
    e = {"dtype": 1, "shape": [3, 4], "shard_id": 0, "offset": 0, "size": 48, "crc32c": ckpt.crc_mask(ckpt.crc32c(a.tobytes()))}
    assert ckpt.verify_entry(e, a.tobytes()) and not ckpt.verify_entry(e, (a + 1).tobytes())


def test_layer_types_match_the_object_graph_of_the_shipped_checkpoint(golden_dir):
    """tests/golden/object_graph_names.json (tests/golden/make_object_graph.py: the TrackableObjectGraph TensorFlow saved
    with the reference's DC2 weights) names the Keras layer behind every checkpoint key: the layer TYPES and their order —
    which weighted layer is a BatchNormalization, a Conv2D, a PReLU, a Dense, a Conv2DTranspose — are the ones the library
    and the oracle implement, and Keras numbered them in that order (conv2d, conv2d_1, ... : creation order = model order)."""
    import json
    import re

    g = json.load(open(os.path.join(golden_dir, "object_graph_names.json")))["full_name"]
    assert set(g) == {k for k, _ in spec.tensor_table()}
    seen = {}
    for prefix, kind in spec.layer_types():
        mine = {k: v for k, v in g.items() if k.rsplit("/", 1)[0] == prefix}
        assert mine, prefix
        layer_names = {v.rsplit("/", 1)[0] for v in mine.values()}
        assert len(layer_names) == 1, (prefix, layer_names)
        name = layer_names.pop()
        m = re.fullmatch(r"(.+?)(?:_(\d+))?", name)
        assert m.group(1) == kind, (prefix, name, kind)
        idx = int(m.group(2) or 0)
        assert idx == seen.get(kind, 0), (prefix, name)  # Keras' per-type counter runs in model order
        seen[kind] = idx + 1
        for k, v in mine.items():  # variable names: kernel / bias / alpha / gamma ... as in the checkpoint key
            assert k.rsplit("/", 1)[1] == v.rsplit("/", 1)[1]
    assert seen == {"batch_normalization": 1, "conv2d": 9, "p_re_lu": 20, "dense": 3, "conv2d_transpose": 8}
