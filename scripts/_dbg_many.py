import sys, threading
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, corpus
from oracle.libzstd_ref import load
from kompressor_b200.mirror import core, io, zstd
libzstd = load("1.5.7")
dictionary = libzstd.train_dictionary([bytes(corpus.generate("logs", 300 + i, 1500)) for i in range(200)], 8192)
raw_dict = b"level=INFO component=scheduler thread=worker-"
jobs = []
for t in range(24):
    kind = ("logs", "json", "literal")[t % 3]
    frames = [bytes(corpus.generate(kind, 1000 + 17 * t + i, 20000 + 3001 * ((t + i) % 5))) for i in range(6)]
    d = (None, dictionary, raw_dict)[t % 4] if t % 4 < 3 else None
    comp = b"".join(libzstd.compress(f, 3, checksum=bool(t & 1), dictionary=d) for f in frames)
    jobs.append(dict(data=b"".join(frames), comp=comp, dict=d, mode="pipe" if t % 2 else "transform"))
for i, j in enumerate(jobs):
    try:
        dec = zstd.ZstdDecompressor(dictionary=j["dict"])
        out = core.transform(dec, j["comp"]) if j["mode"] == "transform" else io.pipe(j["comp"], dec).readByteArray()
        print(i, j["mode"], "dict", None if j["dict"] is None else len(j["dict"]), "ok" if out == j["data"] else "MISMATCH", flush=True)
    except Exception as e:
        print(i, j["mode"], "dict", None if j["dict"] is None else len(j["dict"]), "EXC", e, flush=True)
