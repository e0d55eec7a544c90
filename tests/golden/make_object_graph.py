#!/usr/bin/env python
"""Extracts, from the ONE data shard of the DC2 checkpoint that the reference snapshot contains
(src/debvader/data/weights/dc2/weights_noisy_v4.386--6.61.ckpt.data-00000-of-00002: the TrackableObjectGraph proto TensorFlow
saved next to the weights), the Keras variable names behind every checkpoint key — i.e. the TYPE of every weighted layer of
the reference network in order (batch_normalization, conv2d, p_re_lu, ..., dense, conv2d_transpose, ...) — and commits
them as tests/golden/object_graph_names.json.  Run in the build container (reads /root/reference):

    python tests/golden/make_object_graph.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from debvader_b200.model import ckpt  # noqa: E402
from debvader_b200.model.ckpt import _varint  # noqa: E402

D = "/root/reference/src/debvader/data/weights/dc2"


def fields(buf):
    q = 0
    while q < len(buf):
        tag, q = _varint(buf, q)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, q = _varint(buf, q)
        elif wt == 2:
            ln, q = _varint(buf, q)
            v = buf[q : q + ln]
            q += ln
        elif wt == 5:
            v = buf[q : q + 4]
            q += 4
        elif wt == 1:
            v = buf[q : q + 8]
            q += 8
        else:
            raise ValueError(wt)
        yield f, v


def main():
    prefix = ckpt.latest_checkpoint(D)
    e = ckpt.read_index(prefix + ".index", all_entries=True)["_CHECKPOINTABLE_OBJECT_GRAPH"]
    raw = open(prefix + ".data-00000-of-00002", "rb").read()[e["offset"] : e["offset"] + e["size"]]
    assert ckpt.verify_entry(e, raw), "CRC-32C of the object graph does not match the index"
    ln, p = _varint(raw, 0)
    graph = raw[p + 4 : p + 4 + ln]  # TrackableObjectGraph: nodes = 1 { children = 1, attributes = 2 {name 1, full_name 2, checkpoint_key 3} }
    names = {}
    for f, node in fields(graph):
        if f != 1:
            continue
        for g, attr in fields(node):
            if g != 2:
                continue
            d = dict(fields(attr))
            key, full = d.get(3, b"").decode(), d.get(2, b"").decode()
            if "OPTIMIZER_SLOT" in key or key.startswith("optimizer") or not key.endswith("/.ATTRIBUTES/VARIABLE_VALUE"):
                continue
            names[key[: -len("/.ATTRIBUTES/VARIABLE_VALUE")]] = full
    out = {"source": "TrackableObjectGraph of " + os.path.basename(prefix) + " (reference snapshot, data shard 0; CRC-32C verified)", "full_name": names}
    json.dump(out, open(os.path.join(HERE, "object_graph_names.json"), "w"), indent=1)
    print(len(names), "variables")


if __name__ == "__main__":
    main()
