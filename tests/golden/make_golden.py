#!/usr/bin/env python
"""Generate tests/golden/* by running the REFERENCE's own code (build container only).

Reads /root/reference (read-only) — this script is never run on the GPU box;
its outputs are committed.  tensorflow / sep are not importable here, so they
are stubbed in sys.modules: none of the functions exercised below touches them
(extract_cutouts, DeblendField.get_residual_field / get_predicted_field,
DeblendField.deblend_field with a numpy stand-in for `net`, metrics.mse).

Outputs
  extraction_cases.json   inputs (seeded) + list_idx + sha256 of the cutouts
  field_ops.npz           small residual / predicted-field cases (arrays)
  deblend_field_fake.npz  records of the reference's deblend_field driven by a fake net
  dc2_field2.npz          the packaged field_img_2 + its 40 catalogue centres (BASELINE cfg 0)
  architecture.json       tensor names/shapes parsed from the shipped checkpoint index
"""
import hashlib
import importlib
import json
import os
import struct
import sys
import types

import numpy as np

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    tf = types.ModuleType("tensorflow")
    tf.float32 = np.float32
    tf.cast = lambda x, dt: np.asarray(x, dtype=np.float32)
    sys.modules["tensorflow"] = tf
    sys.modules["sep"] = types.ModuleType("sep")
    sys.path.insert(0, REF)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class _Val:
    def __init__(self, a):
        self._a = a

    def numpy(self):
        return self._a


class FakeDist:
    def __init__(self, mean, std):
        self._m, self._s = mean, std

    def mean(self):
        return _Val(self._m)

    def stddev(self):
        return _Val(self._s)


def fake_net(x):
    """Deterministic stand-in for the VAE: a separable blur + per-band gain, float32."""
    x = np.asarray(x, dtype=np.float32)
    m = (x + np.roll(x, 1, 1) + np.roll(x, -1, 1) + np.roll(x, 1, 2) + np.roll(x, -1, 2)) * np.float32(0.18)
    m = np.maximum(m, 0).astype(np.float32)
    s = (np.float32(1e-4) + np.float32(0.05) * np.abs(x)).astype(np.float32)
    return FakeDist(m, s)


def extraction_cases(extract_cutouts):
    cases = []
    specs = [
        # (seed, F, S, C, centres)
        (1, 15, 5, 3, [[-4, -3], [5, 5], [-5, -5], [6, 6]]),  # the reference's own test values
        (2, 15, 5, 3, [[x, y] for x in range(-20, 21, 1) for y in (-7, 0, 3, 9)]),  # wrap / broadcast sweep
        (3, 64, 9, 2, [[-27.9, 3.2], [27.99, -27.5], [0.5, -0.5], [28, 28], [-28, -28], [29, 0], [0, -29], [1e3, 0]]),
        (4, 259, 59, 6, [[15, 60], [-71, -77], [84, -83], [59, 73], [-11, -24], [10, -34], [87, 16], [42, -97], [-24, -100], [0, 0], [53, 74], [48, 56], [101, 0], [-100, 100], [100, 100], [-101, 3]]),
        (5, 260, 59, 6, [[100, 100], [-100, -100], [101, 101], [-101, -101], [12.7, -12.7]]),
        (6, 33, 59, 6, [[0, 0], [3, 3]]),  # field smaller than the stamp
    ]
    for seed, F, S, C, centres in specs:
        field = np.random.default_rng(seed).random((1, F, F, C))
        cut, idx = extract_cutouts(field.copy(), F, centres, S, C)
        cases.append({"seed": seed, "F": F, "S": S, "C": C, "centres": centres, "list_idx": [int(i) for i in idx], "sha256": sha(cut), "sum": float(cut.sum())})
    return cases


def parse_ckpt_index(path):
    """Minimal reader of a TF tensor-bundle .index (LevelDB table): returns {key: shape}."""
    data = open(path, "rb").read()

    def varint(buf, p):
        r = s = 0
        while True:
            b = buf[p]
            p += 1
            r |= (b & 0x7F) << s
            if not b & 0x80:
                return r, p
            s += 7

    # footer: 48 bytes = metaindex handle + index handle (varints, padded) + 8-byte magic
    footer = data[-48:]
    assert struct.unpack("<Q", footer[-8:])[0] == 0xDB4775248B80FB57
    p = 0
    _, p = varint(footer, p)
    _, p = varint(footer, p)
    ioff, p = varint(footer, p)
    isz, p = varint(footer, p)

    def block_entries(off, sz):
        blk = data[off : off + sz]
        nrestart = struct.unpack("<I", blk[-4:])[0]
        end = len(blk) - 4 - 4 * nrestart
        p, key, out = 0, b"", []
        while p < end:
            shared, p = varint(blk, p)
            non, p = varint(blk, p)
            vlen, p = varint(blk, p)
            key = key[:shared] + blk[p : p + non]
            p += non
            out.append((key, blk[p : p + vlen]))
            p += vlen
        return out

    def parse_entry(v):
        # BundleEntryProto: dtype=1, shape=2 (TensorShapeProto: dim=2 {size=1}), shard_id=3, offset=4, size=5
        p, shape, dtype, size = 0, [], None, None
        while p < len(v):
            tag, p = varint(v, p)
            f, wt = tag >> 3, tag & 7
            if wt == 0:
                val, p = varint(v, p)
                if f == 1:
                    dtype = val
                if f == 5:
                    size = val
            elif wt == 2:
                ln, p = varint(v, p)
                sub = v[p : p + ln]
                p += ln
                if f == 2:
                    q = 0
                    while q < len(sub):
                        t2, q = varint(sub, q)
                        if t2 & 7 == 2:
                            l2, q = varint(sub, q)
                            dim = sub[q : q + l2]
                            q += l2
                            if t2 >> 3 == 2:
                                r = 0
                                while r < len(dim):
                                    t3, r = varint(dim, r)
                                    if t3 & 7 == 0:
                                        val, r = varint(dim, r)
                                        if t3 >> 3 == 1:
                                            shape.append(val)
                                    else:
                                        l3, r = varint(dim, r)
                                        r += l3
                        else:
                            _, q = varint(sub, q)
            elif wt == 5:
                p += 4
            elif wt == 1:
                p += 8
        return dtype, shape, size

    out = {}
    for _, handle in block_entries(ioff, isz):
        boff, q = varint(handle, 0)
        bsz, q = varint(handle, q)
        for key, val in block_entries(boff, bsz):
            k = key.decode()
            if not k or "OPTIMIZER_SLOT" in k or k.startswith("optimizer") or k == "_CHECKPOINTABLE_OBJECT_GRAPH":
                continue
            dtype, shape, size = parse_entry(val)
            out[k.replace("/.ATTRIBUTES/VARIABLE_VALUE", "")] = shape
    return out


def main():
    _stub_modules()
    ext = importlib.import_module("debvader.extract.extraction")
    fd = importlib.import_module("debvader.deblend.field_deblender")
    metrics = importlib.import_module("debvader.training.metrics")
    import pandas as pd

    # ---- extraction ------------------------------------------------------------------
    cases = extraction_cases(ext.extract_cutouts)
    data = "/root/reference/src/debvader/data/dc2_imgs/field/"
    f1 = np.load(data + "field_img.npy")
    offs = [(15, 60), (-71, -77), (84, -83), (59, 73), (-11, -24), (10, -34), (87, 16), (42, -97), (-24, -100), (0, 0), (53, 74), (48, 56)]
    cut, idx = ext.extract_cutouts(f1, 259, offs, 59, 6)
    shipped = np.load(data + "galaxies_from_field.npy")
    pinned = {"galaxies_from_field_equals_reference_extract": bool(np.array_equal(cut, shipped)), "sha256": sha(shipped), "offsets": offs}
    json.dump({"cases": cases, "shipped_fixture": pinned}, open(os.path.join(HERE, "extraction_cases.json"), "w"), indent=1)

    # ---- residual / predicted fields ---------------------------------------------------
    arrays = {}
    for name, F, S, C in (("odd", 45, 9, 2), ("even", 44, 9, 2)):
        rng = np.random.default_rng(100 + F)
        field = rng.normal(0, 1, (1, F, F, C))
        pos = np.array([[0, 0], [3, -5], [-10, 12], [17, 17], [-18, -18], [16, -3], [3, -5], [22, 0]])  # last leaves the canvas partly
        means = rng.random((len(pos), S, S, C)).astype(np.float32)
        stds = rng.random((len(pos), S, S, C)).astype(np.float32)
        rows = {
            "output_images_mean": list(means),
            "output_images_stddev": list(stds),
            "epistemic_uncertainty": list(np.zeros_like(means)),
            "shifts": [np.array([0, 0])] * len(pos),
            "galaxy_distances_to_center_x": list(pos[:, 0]),
            "galaxy_distances_to_center_y": list(pos[:, 1]),
        }
        rec = pd.DataFrame(rows).to_records(index=False)
        obj = fd.DeblendField(None, field, cutout_size=S, nb_of_bands=C)
        arrays[f"{name}_field"] = field
        arrays[f"{name}_pos"] = pos
        arrays[f"{name}_means"] = means
        arrays[f"{name}_stds"] = stds
        arrays[f"{name}_residual"] = obj.get_residual_field(rec)
        pf = obj.get_predicted_field(rec)
        arrays[f"{name}_pred_mean"] = pf["predicted_mean_field"]
        arrays[f"{name}_pred_std"] = pf["predicted_stddev_field"]
        arrays[f"{name}_mse"] = np.array(metrics.mse(field, arrays[f"{name}_residual"]))
    np.savez_compressed(os.path.join(HERE, "field_ops.npz"), **arrays)

    # ---- deblend_field driven by a fake net --------------------------------------------
    F, S, C = 101, 59, 6
    rng = np.random.default_rng(7)
    field = rng.normal(0, 0.5, (1, F, F, C))
    field[0, 40:60, 45:58, :] += 30.0
    centres = np.array([[0.0, 0.0], [10.0, -12.0], [40.0, 0.0], [-21.0, 21.0], [21.0, 21.0], [-22.0, 0.0], [5.0, 5.0]])
    obj = fd.DeblendField(fake_net, field, cutout_size=S, nb_of_bands=C)
    rec = obj.deblend_field(centres, mse_criterion=2.0)
    out = {
        "field": field,
        "centres": centres,
        "list_idx": np.array(list(rec["list_idx"]), dtype=np.int64),
        "passed_cuts": np.array(list(rec["passed_cuts"]), dtype=bool),
        "mean": np.stack(list(rec["output_images_mean"])),
        "stddev": np.stack(list(rec["output_images_stddev"])),
        "cutouts": np.stack(list(rec["cutout_images"])),
        "dx": np.array(list(rec["galaxy_distances_to_center_x"])),
        "dy": np.array(list(rec["galaxy_distances_to_center_y"])),
        "residual": obj.get_residual_field(),
        "nb_detected": np.array(obj.nb_of_detected_objects),
        "nb_deblended": np.array(obj.nb_of_deblended_galaxies),
        "record_names": np.array(rec.dtype.names),
    }
    np.savez_compressed(os.path.join(HERE, "deblend_field_fake.npz"), **out)

    # ---- BASELINE cfg 0 inputs ----------------------------------------------------------
    f2 = np.load(data + "field_img_2.npy")
    cat = np.load(data + "gal_coordinates_complete_truth_catalog_2.npy")
    cen = np.load(data + "field_center_2.npy")
    centres2 = np.round(np.stack([cat[:, 1] - cen[1], cat[:, 0] - cen[0]], axis=1))
    cut2, idx2 = ext.extract_cutouts(f2, 259, centres2, 59, 6)
    np.savez_compressed(
        os.path.join(HERE, "dc2_field2.npz"),
        field=f2,
        centres=centres2,
        list_idx=np.array(idx2, dtype=np.int64),
        cutouts_sha256=np.array(sha(cut2)),
        stamps=np.load("/root/reference/src/debvader/data/dc2_imgs/imgs_dc2.npy")[:4].astype(np.float32),
    )

    # ---- architecture -------------------------------------------------------------------
    shapes = parse_ckpt_index("/root/reference/src/debvader/data/weights/dc2/weights_noisy_v4.386--6.61.ckpt.index")
    json.dump({"source": "weights/dc2/weights_noisy_v4.386--6.61.ckpt.index", "net_summary_params": {"encoder": 3741224, "decoder": 4577228, "total": 8318452}, "tensors": shapes}, open(os.path.join(HERE, "architecture.json"), "w"), indent=1, sort_keys=True)
    print("golden written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
