import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def gbytes(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def expected_of(v):
    """(expected bytes or None, sha256) of a manifest vector."""
    p = os.path.join(GOLDEN, v["name"] + ".raw")
    return (gbytes(v["name"] + ".raw") if os.path.exists(p) else None), v["sha256"]


def check_output(v, out: bytes):
    exp, sha = expected_of(v)
    assert len(out) == v["size"], v["name"]
    if exp is not None:
        assert out == exp, v["name"]
    assert hashlib.sha256(out).hexdigest() == sha, v["name"]


def frames_of(c):
    return [bytes(c.comp[int(c.offsets[i]):int(c.offsets[i] + c.sizes[i])]) for i in range(c.n_frames)]


def table_of(frames):
    n = len(frames)
    blob = b"".join(frames)
    comp = np.frombuffer(blob, dtype=np.uint8).copy() if blob else np.zeros(1, np.uint8)
    sizes = np.array([len(f) for f in frames], dtype=np.uint64)
    offs = np.zeros(n, np.uint64)
    if n:
        offs[1:] = np.cumsum(sizes)[:-1]
    return comp, offs, sizes
