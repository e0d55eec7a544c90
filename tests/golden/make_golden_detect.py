#!/usr/bin/env python
"""tests/golden/detect_dc2.npz: detections of the packaged DC2 field (dc2_field2.npz = the reference's field_img_2.npy) by
oracle/detect_numpy.py.  `sep` (what the reference calls, detect/detection.py:15,37) is not installable here, so this golden
pins the ORACLE against drift — not sep.  Where sep is available, run with --sep to store its detections next to the oracle's
(`sep_centres`): that would pin the order contract."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from debvader_b200.detect.detection import FILTER_KERNEL, detect_objects  # noqa: E402
from oracle import detect_numpy as D  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
field = np.load(os.path.join(here, "dc2_field2.npz"))["field"]
c, dt = D.detect(field, FILTER_KERNEL, return_details=True)
out = {"centres": c, "last": dt["last"], "x": dt["x"], "y": dt["y"], "npix": dt["npix"], "globalback": dt["globalback"], "globalrms": dt["globalrms"],
       "thresh": dt["thresh"]}
if "--sep" in sys.argv:
    out["sep_centres"] = detect_objects(field, backend="sep")
np.savez_compressed(os.path.join(here, "detect_dc2.npz"), **out)
print(len(c), "detections; global rms", dt["globalrms"])
