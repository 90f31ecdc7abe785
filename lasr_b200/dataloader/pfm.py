"""PFM reader with the reference's signature and conventions (third_party/ext_utils/util_flow.py:36-80):
returns ``(data, scale)``, rows flipped to top-down, 3 channels for 'PF' and 1 for 'Pf'; a negative scale in the
header means little-endian.  Host-side file decode; the arithmetic that follows runs on the device."""
from __future__ import annotations

import re

import numpy as np


def readPFM(file):
    with open(file, "rb") as fh:
        header = fh.readline().rstrip().decode("utf-8")
        if header not in ("PF", "Pf"):
            raise Exception("Not a PFM file.")
        dims = re.match(r"^(\d+)\s(\d+)\s$", fh.readline().decode("utf-8"))
        if not dims:
            raise Exception("Malformed PFM header.")
        width, height = (int(v) for v in dims.groups())
        scale = float(fh.readline().rstrip().decode("utf-8"))
        endian = "<" if scale < 0 else ">"
        data = np.fromfile(fh, endian + "f")
    shape = (height, width, 3) if header == "PF" else (height, width)
    return np.flipud(np.reshape(data, shape)), abs(scale)
