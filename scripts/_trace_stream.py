import ctypes, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle.libzstd_ref import load
from kompressor_b200._lib import lib
from test_gpu_stream import _InBuf, _OutBuf
z = load("1.5.7"); K = lib(); Z = z.lib
data = open(sys.argv[1], "rb").read(); out_slice = int(sys.argv[2]); in_chunk = int(sys.argv[3]) if sys.argv[3] != "None" else None
def trace(L, pre):
    getattr(L, pre + "createDCtx").restype = ctypes.c_void_p
    getattr(L, pre + "freeDCtx").argtypes = [ctypes.c_void_p]
    f = getattr(L, pre + "decompressStream"); f.argtypes = [ctypes.c_void_p] * 3; f.restype = ctypes.c_size_t
    ise = getattr(L, pre + "isError"); ise.argtypes = [ctypes.c_size_t]
    d = getattr(L, pre + "createDCtx")()
    src = ctypes.create_string_buffer(bytes(data), len(data)); base = ctypes.addressof(src)
    visible, pos = (len(data) if in_chunk is None else 0), 0
    dst = ctypes.create_string_buffer(out_slice); wpos = 0; total = 0; rows = []
    for _ in range(100000):
        if in_chunk is not None and pos == visible and visible < len(data): visible = min(len(data), visible + in_chunk)
        inb = _InBuf(base, visible, pos); outb = _OutBuf(ctypes.addressof(dst), out_slice, wpos)
        r = f(d, ctypes.byref(outb), ctypes.byref(inb))
        if ise(r): rows.append((visible, pos, inb.pos, wpos, outb.pos, "ERR %d" % ((1 << 64) - r))); break
        total += outb.pos - wpos
        rows.append((visible, pos, inb.pos, wpos, outb.pos, r, total))
        pos, wpos = inb.pos, outb.pos
        ins = wpos == out_slice and r != 0
        if wpos == out_slice: wpos = 0
        if pos == visible and not ins and visible == len(data): break
    return rows
a, b = trace(Z, "ZSTD_"), trace(K, "kzg_")
print("call: (visible, pos_in, pos_out | wpos_in, wpos_out, result, total)   libzstd | kzg")
for i in range(max(len(a), len(b))):
    x = a[i] if i < len(a) else None; y = b[i] if i < len(b) else None
    strip = lambda t: None if t is None else (t[:5] + ((t[5] == 0) if not isinstance(t[5], str) else t[5],) + t[6:])
    mark = "" if strip(x) == strip(y) else "   <<<"
    print(i, x, "|", y, mark)
