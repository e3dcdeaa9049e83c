"""Minimal NRRD writer / reader for decode_embeddings.py (the reference calls pynrrd's `nrrd.write(path, data,
header={'spacings': ...})`, decode_embeddings.py:50; pynrrd's default index order is Fortran: `sizes` = data.shape
and the first axis varies fastest in the payload).  Raw encoding, little endian."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

_TYPES = {"int64": "int64", "int32": "int32", "int16": "int16", "uint8": "uint8", "float32": "float", "float64": "double"}
_BACK = {v: k for k, v in _TYPES.items()}


def write_nrrd(path: str, data: np.ndarray, header: Dict = None) -> None:
    if data.dtype.name not in _TYPES:
        raise TypeError(f"write_nrrd: unsupported dtype {data.dtype}")
    lines = ["NRRD0004", f"type: {_TYPES[data.dtype.name]}", f"dimension: {data.ndim}",
             "sizes: " + " ".join(str(s) for s in data.shape)]
    for k, v in (header or {}).items():
        lines.append(f"{k}: " + (" ".join(str(x) for x in v) if isinstance(v, (tuple, list)) else str(v)))
    lines += ["endian: little", "encoding: raw", "", ""]
    with open(path, "wb") as f:
        f.write("\n".join(lines).encode("ascii"))
        f.write(np.asfortranarray(data).astype(data.dtype.newbyteorder("<"), copy=False).tobytes(order="F"))


def read_nrrd(path: str) -> Tuple[np.ndarray, Dict[str, str]]:
    raw = open(path, "rb").read()
    head, payload = raw.split(b"\n\n", 1)
    fields = dict(l.split(": ", 1) for l in head.decode("ascii").splitlines()[1:] if ": " in l)
    shape = tuple(int(s) for s in fields["sizes"].split())
    data = np.frombuffer(payload, dtype=np.dtype(_BACK[fields["type"]]).newbyteorder("<")).reshape(shape, order="F")
    return data, fields
